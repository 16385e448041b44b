/*
 * ionization_b200 -- C-ABI of the B200-native mesh time-evolution engine.
 *
 * This is the drop-in boundary for the hot path of JoshKarpel/ionization
 * (reference citations are relative to /root/reference/ionization):
 *
 *   MeshSimulation.run()            mesh/sims.py:255-346
 *     -> QuantumMesh.evolve()       mesh/meshes.py:251-257
 *       -> EvolutionMethod.evolve() mesh/evolution_methods.py:19-24
 *         -> DotOperator/TDMAOperator._apply   mesh/mesh_operators.py:96-111
 *           -> cy.tdma()            cy.pyx:9-50          (the reference's only native symbol)
 *     -> Datastore.store()          mesh/data.py:189-464 (norm, inner products, ...)
 *
 * Plain pointers and sizes only; every complex array is interleaved (re, im)
 * float64 ("complex128"), i.e. exactly numpy's / C99's `double complex` layout.
 * All functions return 0 on success and a negative ION_E* code on failure;
 * ion_last_error() returns a thread-local human-readable message.  The
 * library never computes on the CPU: without a CUDA device every compute
 * entry point fails with ION_ENODEVICE.
 *
 * Unless stated otherwise pointer arguments are HOST pointers; the library
 * performs the host<->device copies on the simulation's stream.
 */
#ifndef IONIZATION_B200_H
#define IONIZATION_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ION_ABI_VERSION 2

/* status codes */
#define ION_OK 0
#define ION_EINVAL (-1)    /* bad argument / inconsistent shapes            */
#define ION_ENODEVICE (-2) /* no usable CUDA device                         */
#define ION_ECUDA (-3)     /* CUDA runtime error (see ion_last_error)       */
#define ION_ESTATE (-4)    /* call order violated (e.g. step before set_*)  */
#define ION_ENOTSUP (-5)   /* configuration not supported by this build     */

/* evolution programs: (mesh, gauge, method) triples of the reference */
#define ION_SH_LEN_SO 0   /* SphericalHarmonicLengthGaugeOperators + SplitInteractionOperator  (mesh_operators.py:815-1127, evolution_methods.py:80-123) */
#define ION_SH_VEL_SO 1   /* SphericalHarmonicVelocityGaugeOperators + SplitInteractionOperator (mesh_operators.py:1130-1408)                           */
#define ION_LINE_LEN_CN 2 /* LineLengthGaugeOperators + AlternatingDirectionImplicit            (mesh_operators.py:304-349, evolution_methods.py:46-77) */
#define ION_LINE_LEN_SO 3 /* LineLengthGaugeOperators + SplitInteractionOperator                (mesh_operators.py:329-341)                              */
#define ION_LINE_VEL_SO 4 /* LineVelocityGaugeOperators + SplitInteractionOperator              (mesh_operators.py:352-427)                              */
#define ION_SH_LEN_ADI 5  /* SphericalHarmonicLengthGaugeOperators + AlternatingDirectionImplicit (evolution_methods.py:46-77, mesh_operators.py:1020-1035); unsharded, l_bound <= 4096 */

/* observables computed by ion_sim_observe / ion_sim_run (bit mask) */
#define ION_OBS_NORM 1u            /* QuantumMesh.norm                    mesh/meshes.py:215-217          */
#define ION_OBS_INNER_PRODUCTS 2u  /* inner_product(state)                mesh/meshes.py:1099-1131        */
#define ION_OBS_NORM_BY_L 4u       /* SphericalHarmonicMesh.norm_by_l     mesh/meshes.py:1133-1136        */
#define ION_OBS_R 8u               /* r_expectation_value                 mesh/meshes.py:235-237          */
#define ION_OBS_Z 16u              /* z_expectation_value                 mesh/meshes.py:231-233          */
#define ION_OBS_H0 32u             /* internal_energy_expectation_value   mesh/meshes.py:219-223          */
#define ION_OBS_NORM_WITHIN 64u    /* NormWithinRadius.store              mesh/data.py:419-422            */
#define ION_OBS_HINT 128u          /* <H_int/field>: total energy = H0 + field*HINT (mesh/meshes.py:225-229) */

typedef struct ion_sim ion_sim_t;

/* ---- library ---------------------------------------------------------------- */
int ion_abi_version(void);
const char *ion_last_error(void);
int ion_device_count(void);

/* ---- (1) fine-grained entry point: replaces cy.tdma (cy.pyx:9) ---------------
 * Solves `batch` independent tridiagonal systems M x = rhs without pivoting
 * (the reference never pivots, cy.pyx:34-48).  Arrays are [batch][n-1] (sub,
 * sup) and [batch][n] (diag, rhs, x); sub[i] is M[i+1][i], sup[i] is M[i][i+1],
 * i.e. scipy dia_matrix data[0][:-1] and data[2][1:] (cy.pyx:20-22).
 * x may alias rhs.  Inputs are not modified (as cy.tdma). */
int ion_tdma_c128(const void *sub, const void *diag, const void *sup, const void *rhs, void *x, int64_t n,
                  int64_t batch, int device);

/* ---- (2) device-resident simulation: what MeshSimulation.run() needs ---------
 * One handle = `batch` independent simulations on ONE mesh sharing every
 * coefficient vector (a scan ensemble, ionization_scans/scan_mesh.py:40-68;
 * batch = 1 is a plain simulation).  g is [batch][L][R] complex128, r fastest
 * -- the reference's own storage (mesh/meshes.py:1003, :1052-1064); L = 1 for
 * LineMesh (R = z_points).
 *
 * l-block sharding (one large simulation on several GPUs): the handle owns the
 * channels [l_begin, l_begin + L) of a mesh with L_total channels; see
 * ion_sim_halo_* below.  For an unsharded simulation l_begin = 0, L_total = L. */
int ion_sim_create(int program, int64_t L, int64_t R, int64_t batch, int device, ion_sim_t **out);
int ion_sim_create_sharded(int program, int64_t L_total, int64_t l_begin, int64_t L, int64_t R, int64_t batch,
                           int device, ion_sim_t **out);
int ion_sim_destroy(ion_sim_t *sim);

/* Launch on an existing CUDA stream (a cudaStream_t, e.g. torch.cuda.current_stream().cuda_stream).
 * Default: the legacy default stream. */
int ion_sim_set_stream(ion_sim_t *sim, void *cuda_stream);

/* Field-free Hamiltonian, tridiagonal in r per channel (mesh_operators.py:889-928, :244-269, :310-318):
 * h_diag complex128 [L][R] (owned channels only), h_off float64 [R-1] (coupling j<->j+1, the same for every
 * channel).  Crank-Nicolson uses (1 -+ i tau H0) (evolution_methods.py:98-111). */
int ion_sim_set_hamiltonian(ion_sim_t *sim, const void *h_diag, const double *h_off);

/* Length gauge, SphericalHarmonicMesh (mesh_operators.py:988-1080): angle(l,j) = s * c_l[l] * x_j[j] with
 * s = tau*E per step, c_l float64 [L_total-1] (pair l<->l+1, GLOBAL l), x_j = -q r_j float64 [R]. */
int ion_sim_set_len_coupling(ion_sim_t *sim, const double *c_l, const double *x_j);

/* Velocity gauge, SphericalHarmonicMesh (mesh_operators.py:1143-1408): theta1 = s * f1_l[l] * y_j[j],
 * theta2 = s * c_l[l] * z_j[j], s = tau*A per step.  c_l, f1_l: [L_total-1]; y_j: [R]; z_j: [R-1]. */
int ion_sim_set_vel_coupling(ion_sim_t *sim, const double *c_l, const double *f1_l, const double *y_j,
                             const double *z_j);

/* LineMesh (mesh_operators.py:320-341, :358-427): w_z = -q z float64 [R] (length gauge; may be NULL for
 * ION_LINE_VEL_SO), v_pref = hbar (q/m) / (2 dz) (velocity gauge). */
int ion_sim_set_line_coupling(ion_sim_t *sim, const double *w_z, double v_pref);

/* Radial mask applied after every step (mesh/meshes.py:257, potentials/masks.py:76-89): float64 [R] or NULL. */
int ion_sim_set_mask(ion_sim_t *sim, const double *mask);

/* Observables set-up: inner-product multiplier (delta_r / delta_z, mesh/meshes.py:1012, :296), coordinate
 * vector r_j [R] (for <r>, <z>, norm-within-radius), test states (n rows complex128 [n][R]; state_l[n] = GLOBAL
 * channel of each row, 0 for LineMesh) and radii (float64 [n_radii]).  Any pointer may be NULL / count 0. */
int ion_sim_set_observables(ion_sim_t *sim, double inner_product_multiplier, const double *r_j, int64_t n_states,
                            const int64_t *state_l, const void *state_rows, int64_t n_radii, const double *radii);

/* g <-> host, reference layout [batch][L][R]. */
int ion_sim_write_g(ion_sim_t *sim, const void *g);
int ion_sim_read_g(ion_sim_t *sim, void *g);
/* the same initial state for every member of an ensemble (ionization_scans/scan_mesh.py:40-68 gives every member the
 * same initial_state): g is ONE member, [L][R]; it is copied to the device once and replicated there. */
int ion_sim_write_g_broadcast(ion_sim_t *sim, const void *g);

/* Advance n_steps time steps (each = QuantumMesh.evolve(): evolution operators then mask).
 *   taus   float64 [n_steps]          tau_n = (t_n - t_{n-1}) / (2 hbar)   (evolution_methods.py:92)
 *   fields float64 [n_steps][batch]   the field scalar the reference samples for that step:
 *                                     E(t_n + dt/2) for SH length gauge (mesh_operators.py:1011-1013),
 *                                     E(t_n) for Line length gauge (:321-323), A(t_0..t_n) for velocity
 *                                     gauge (:1184-1186, :373-375).
 * Returns after the work is ENQUEUED on the stream (asynchronous); use ion_sim_synchronize(). */
int ion_sim_step(ion_sim_t *sim, int64_t n_steps, const double *taus, const double *fields);

/* Size in float64 of one observation record per simulation for the given mask:
 * layout [norm][ip re,im x n_states][norm_by_l x L][<r>][<z>][<H0>][<Hint/field>][norm_within x n_radii], each
 * present only if its bit is set, in this order. */
int64_t ion_sim_observation_size(ion_sim_t *sim, uint32_t what);

/* Observe the current state: out float64 [batch][ion_sim_observation_size].  Synchronous. */
int ion_sim_observe(ion_sim_t *sim, uint32_t what, double *out);

/* The device-resident loop of MeshSimulation.run() (mesh/sims.py:289-321): advance n_steps steps and record an
 * observation after every step n (0-based) with observe_mask[n] != 0 (NULL: none).  out receives the records
 * in step order: [n_observed][batch][observation_size].  Synchronous. */
int ion_sim_run(ion_sim_t *sim, int64_t n_steps, const double *taus, const double *fields,
                const uint8_t *observe_mask, uint32_t what, double *out);

int ion_sim_synchronize(ion_sim_t *sim);

/* ---- device-side access for the multi-GPU plumbing (torch.distributed / NCCL) ----
 * l-block sharding: the rotation sweeps couple the last owned channel with the first channel of the next
 * shard.  ion_sim_step_phase() runs one step in phases so the caller can exchange boundary channels between
 * them; the boundary buffers are device memory, [batch][R_padded] complex128 each, in the engine's internal
 * row order (opaque, identical on every shard of the same mesh).
 *   which: 0 = send-to-lower (my first channel), 1 = send-to-upper (my last channel),
 *          2 = recv-from-lower,                   3 = recv-from-upper. */
int ion_sim_halo_buffer(ion_sim_t *sim, int which, void **device_ptr, int64_t *n_bytes);
/* number of phases (pair-local kernels) of one step of this program, and whether the neighbours' boundary channels
 * must be delivered into this shard's recv buffers before a given phase.  Shards are cut at even channels, so only
 * the odd-parity sweeps cross a cut: 1 exchange per step in the length gauge, 3 in the velocity gauge.
 * A shard holds one ghost channel per neighbour: ion_sim_set_hamiltonian() of a shard takes h_diag for the channels
 * [l_begin - (l_begin > 0), l_begin + L + (l_begin + L < L_total)), ghosts included. */
int ion_sim_num_phases(ion_sim_t *sim);
int ion_sim_phase_needs_halo(ion_sim_t *sim, int phase);
/* run phase `phase` of the step with scalars tau, field[batch] (host pointers).  Before phase p>0 the caller
 * must have delivered the neighbours' send buffers (filled by phase p-1) into this shard's recv buffers. */
int ion_sim_step_phase(ion_sim_t *sim, int phase, double tau, const double *field);

/* ---- halo exchange over NVLink peer memory, inside the engine (csrc/halo.cuh) ----
 * Instead of the caller exchanging boundary channels between phases, the shards can be linked once: every shard
 * exports a blob (CUDA IPC handle of its halo block: flags + two staging slots per side), the caller hands
 * each shard the blobs of its neighbours (torch.distributed all_gather of ION_PEER_BLOB_BYTES bytes is the only
 * collective), and from then on ion_sim_step / ion_sim_run advance the shard like an unsharded simulation: before every
 * odd-parity kernel a small kernel stores the boundary channel straight into the neighbour's staging slot, raises a
 * flag there, waits for its own flag and moves the received channel into its ghost channel.  Every shard must make the same sequence of calls (the exchange is a
 * rendezvous).  Replaces the cgo/NCCL-style send/recv the reference would need; the reference itself has no multi-device
 * path.  same_process != 0: the blob's raw pointers are used (several shards in one process, tests). */
#define ION_PEER_BLOB_BYTES 96
int ion_sim_export_peer(ion_sim_t *sim, void *blob, int64_t blob_bytes);
int ion_sim_attach_peer(ion_sim_t *sim, int side /* 0: lower neighbour, 1: upper */, const void *blob, int64_t blob_bytes,
                        int same_process);
/* one exchange outside a step (e.g. before observing <z>, which couples the last owned channel to the upper ghost) */
int ion_sim_exchange_halos(ion_sim_t *sim);
/* build the LU factors for tau now (everything that allocates), so that later steps only launch kernels */
int ion_sim_prepare(ion_sim_t *sim, double tau);
/* size every buffer a later ion_sim_step / ion_sim_run of n_steps steps with n_records observations of `what` needs, now.
 * Linked shards must not allocate or free between hand-shakes (cudaFree waits for the device to go idle, which it never does
 * while a neighbour's exchange kernel is waiting for this shard): call ion_sim_prepare and ion_sim_reserve on every shard first. */
int ion_sim_reserve(ion_sim_t *sim, int64_t n_steps, int64_t n_records, uint32_t what);
/* exchanges completed so far; aborted != 0: a hand-shake timed out (a neighbour never arrived), the state is invalid */
int ion_sim_halo_status(ion_sim_t *sim, int64_t *exchanges_done, int *aborted);

/* raw device pointer of the wavefunction in internal layout + its size (for peer copies / checksums) */
int ion_sim_device_psi(ion_sim_t *sim, void **device_ptr, int64_t *n_bytes);

/* ---- field set-up on the device -------------------------------
 * The per-step field scalars of a whole scan of windowed Sinc pulses (potentials/pulses.py:599-1013, windows.py:129-168) in the
 * layout ion_sim_step takes: out float64 [n_times - 1][n_pulses].
 *   kind ION_FIELD_E: out[n-1][b] = E_b(times[n] + t_offset)                                   (length gauge: mesh_operators.py:1011-1013, :321-323)
 *   kind ION_FIELD_A: out[n-1][b] = -simps(E_b(times[0..n]), times[0..n]), old-scipy even='avg' (velocity gauge: :1184-1186, pulses.py:58-77),
 *                     every prefix in one O(n) sweep per pulse instead of the reference's O(n^2).
 * pulse_params float64 [n_pulses][8]: amplitude, delta_omega, omega_carrier, phase, pulse_center, window_time, window_width
 * (<= 0: no window), window_center. */
#define ION_FIELD_E 0
#define ION_FIELD_A 1
int ion_sinc_pulse_fields(int device, int kind, int64_t n_times, const double *times, double t_offset, int64_t n_pulses,
                          const double *pulse_params, double *out);

/* ---- measurement ------------------------------------------------------------- */
/* Number of kernel launches issued by this handle since creation (bench.py's gpu_launches). */
int64_t ion_sim_launch_count(ion_sim_t *sim);
/* Profile: run n_steps steps with CUDA events around every kernel launch; returns, per kernel kind
 * (ion_kernel_name(k)), accumulated milliseconds and launch counts.  Arrays of length ion_num_kernel_kinds(). */
int ion_num_kernel_kinds(void);
const char *ion_kernel_name(int kind);
int ion_sim_profile(ion_sim_t *sim, int64_t n_steps, const double *taus, const double *fields, double *ms,
                    int64_t *launches);
/* Measured FP64 pipe peak of the device: thread-level double-precision FMAs per second of independent chains at full
 * occupancy (x2 = FLOP/s).  The second bound bench.py reports next to the HBM roofline. */
int ion_fp64_peak(int device, double *fma_per_second);

#ifdef __cplusplus
}
#endif
#endif /* IONIZATION_B200_H */
