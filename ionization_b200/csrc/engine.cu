// ionization_b200 -- host side of the engine and the C-ABI (include/ionization_b200.h).
//
// One `ion_sim` = `batch` simulations on one mesh, wavefunction resident in HBM in the row-interleaved
// layout of kernels.cuh.  ion_sim_step()/ion_sim_run() translate the reference's operator sequence
// (evolution_methods.py:89-123, SURVEY.md App. C "operator order per step") into pair-local kernels:
//
//   length gauge   E_e E_o CN E_o E_e mask      ->  [ROT even(s_n [+ s_n+1])] [ROT-CN-ROT odd]            2 launches/step
//   velocity gauge h1_e h1_o h2_ee h2_eo h2_oe h2_oo CN (reversed) mask
//                                                ->  [ROT odd] [H2 even] [H2-CN-H2 odd] [H2 even rev] [ROT odd] [ROT even]
//
// Consecutive operators acting on the same l-pairs are fused into one kernel; the trailing even rotation
// of step n, the mask and the leading even rotation of step n+1 commute with each other (all diagonal in r
// on the same pair) and are fused across the step boundary whenever no observation is requested in between.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/ionization_b200.h"
#include "kernels.cuh"
#include "slab.cuh"
#include "adi.cuh"
#include "ensemble.cuh"
#include "halo.cuh"
#include "fields.cuh"

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string &msg)
{
    g_last_error = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                                              \
    do {                                                                                                            \
        cudaError_t _e = (expr);                                                                                    \
        if (_e != cudaSuccess)                                                                                      \
            return fail(ION_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" __FILE__ ":" +         \
                                       std::to_string(__LINE__) + ")");                                             \
    } while (0)

enum KernelKind : int {
    KK_ROT = 0,
    KK_ROT_CN_ROT,
    KK_H2,
    KK_H2_CN_H2,
    KK_CN,
    KK_LINE_SO_LEN,
    KK_LINE_SO_VEL,
    KK_LINE_CN,
    KK_SWEEP_FLAT,
    KK_MASK,
    KK_OBSERVE,
    KK_SLAB,
    KK_LEN_STEP,
    KK_HALO,
    KK_ADI_L,
    KK_LEN_ENS,
    KK_COUNT
};
const char *const kKernelNames[KK_COUNT] = {"rot",        "rot_cn_rot", "h2",        "h2_cn_h2", "cn",     "line_so_len",
                                            "line_so_vel", "line_cn",    "sweep_flat", "mask",    "observe",
                                            "slab",        "len_step",   "halo",
                                            "adi_l",       "len_ens"};

template <typename T>
int dev_alloc(T **p, size_t n)
{
    if (*p) cudaFree(*p);
    *p = nullptr;
    if (n == 0) return ION_OK;
    CUDA_TRY(cudaMalloc((void **)p, n * sizeof(T)));
    return ION_OK;
}

}  // namespace



struct ion_sim {
    int program = 0;
    // L / l_begin describe the channels HELD in psi: the owned ones plus one ghost channel on either side of an
    // l-block shard (g_lo, g_hi); L_own / l_own are the owned range.  Unsharded: L == L_own == L_total.
    int L = 0, R = 0, batch = 0, device = 0, L_total = 0, l_begin = 0;
    int L_own = 0, l_own = 0, g_lo = 0, g_hi = 0;
    // parity of the channels an l-block shard is cut at.  0: blocks begin on even channels -- odd-parity pairs straddle the cuts
    // (every program).  1: blocks begin on odd channels (length gauge) -- every odd-parity pair is local, so the folded one-pass step
    // (PROG_LEN_STEP) runs on shards too: its read-only even-pair partners are the ghost channels.
    int cut_parity = 0;
    // T = threads per channel (row stride of the layout); a channel is cut into S r-segments of T_seg interior threads,
    // each CTA running Tc = T_seg + 2H threads (S == 1: Tc == T_seg == T, H == 0)
    int M = 4, T = 0, Rp = 0, tmax = 0, S = 1, T_seg = 0, H = 0, Tc = 0;
    double *scal_phase = nullptr;
    bool line = false;
    cudaStream_t stream = 0;

    cplx *psi = nullptr, *io_stage = nullptr;
    // second wavefunction buffer of the out-of-place inter-solve kernel (slab.cuh); psi always names the buffer that
    // holds the current state, psi_home the one it must be in whenever control returns to the caller
    cplx *psi2 = nullptr, *psi_home = nullptr;
    bool use_slab = true;
    int slab_state = 0;  // 0: not examined, 1: usable, -1: not usable
    // l-block shards: halo exchange over peer memory (halo.cuh)
    unsigned long long *hflags = nullptr;           // my halo block: HF_COUNT flags, then staging[2 sides][2 slots][Rp]
    unsigned long long *peer_flags[2] = {nullptr, nullptr};
    cplx *peer_stage[2] = {nullptr, nullptr};       // the neighbours' staging slots facing this shard (peer mappings)
    // exchange fused into the folded length-gauge step (kernels.cuh: PROG_LEN_STEP_HALO): the neighbours' fused staging slots
    // facing this shard, and the per-(side, segment) launch counters of the boundary CTAs
    cplx *peer_fstage[2] = {nullptr, nullptr};
    unsigned long long *hf_sent = nullptr;
    bool use_fused_halo = true;
    void *peer_ipc_base[2] = {nullptr, nullptr};    // IPC mappings to close
    bool peers_attached = false;
    bool neighbour_on_same_device = false;  // a linked neighbour lives on this GPU (several shards of one process on one device)
    bool use_len_fold = true;
    bool use_ens = true;
    int ens_state = 0;  // scan ensembles: persistent folded length-gauge kernel with a prefetch pipeline (ensemble.cuh); 0 / 1 / -1 as above
    int ens_ctas = 0;
    int len_fold_state = 0;  // length gauge: even sweep folded into the out-of-place PROG_LEN_STEP kernel (0 / 1 / -1 as above)
    int slab_G = 0, slab_slabs = 0, slab_chunks = 0, slab_Qc = 0, slab_nQ = 0, slab_threads = 0;
    cplx *h_diag = nullptr;
    double *h_off = nullptr;
    std::vector<double> h_off_host;
    cplx *w = nullptr, *aggP = nullptr, *aggQ = nullptr, *th = nullptr;
    cplx *thd = nullptr;  // [L][M][T] tau * h_diag, permuted (ION_SH_LEN_ADI: explicit r half-step, adi.cuh)
    double *toff = nullptr, *toff_prev = nullptr;
    double *vec = nullptr, *zvec = nullptr, *zprev = nullptr, *mask = nullptr, *rvec = nullptr;
    double *cl = nullptr, *cl2 = nullptr, *cl_z = nullptr;
    double *scal = nullptr;
    size_t scal_cap = 0;

    double ipm = 1.0;
    int n_states = 0, n_radii = 0;
    double radii[ION_MAX_RADII] = {0};
    cplx *state_rows = nullptr;
    int *state_first = nullptr, *state_order = nullptr;
    double *partial = nullptr, *ip_out = nullptr, *obs_out = nullptr;
    // length gauge, fused observation: second set of the two buffers above -- the record of step n is assembled on the side branch while
    // the kernel of step n + 1 already fills the other set (obs_parity, ev_done as for the slab kernel)
    double *partial2 = nullptr, *ip_out2 = nullptr;
    // velocity gauge, fused observation by STORING the observed state (slab.cuh: k_slab<true, true>): two buffers of psi's size, reduced by
    // k_observe on the side branch
    cplx *obs_psi[2] = {nullptr, nullptr};
    double *slab_partial = nullptr, *slab_ip = nullptr;  // per-slab partial sums of the fused observation (slab.cuh)
    unsigned *obs_counter = nullptr;                     // [batch] CTAs of k_slab_obs_assemble that are done
    // the record assembly of a fused observation runs on a side branch of the captured graph, so that the next step's kernels
    // depend on k_slab only: slab_partial / slab_ip are double-buffered (obs_parity), ev_fork orders assemble after its k_slab,
    // ev_done[k] lets the k_slab that reuses buffer k (two observations later) and the end of a chunk wait for the assembly
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_done[2] = {nullptr, nullptr};
    bool side_pending[2] = {false, false};
    int obs_parity = 0;
    size_t slab_partial_half = 0, slab_ip_half = 0;
    size_t obs_cap = 0;

    double vec_dv = 0.0;  // length gauge: increment of the coupling vector per radial row when it is linear to rounding, else 0
    bool have_h = false, have_coupling = false;
    double factored_tau = 0.0;
    bool factored = false;
    int short_scan = 0;  // reach of the cross-warp scan inflow in warps; 0 = full scan
    int64_t launch_count = 0;

    // CUDA graphs: chunks of up to GRAPH_CHUNK consecutive steps are captured once and replayed; the kernels of a
    // captured chunk read their per-step scalars from scal_chunk and write observations to obs_chunk, which are
    // refilled / drained around every replay with device-to-device copies.
    struct GraphEntry {
        cudaGraphExec_t exec = nullptr;
        int64_t launches = 0;
    };
    std::map<std::string, GraphEntry> graphs;
    double *scal_chunk = nullptr, *obs_chunk = nullptr;
    size_t obs_chunk_cap = 0;
    bool own_stream = false;
    bool use_graphs = true;
    bool use_pdl = true;
    bool capturing = false;
    // per-call scalars: two pinned staging slots, so that ion_sim_step returns without waiting for the stream
    double *scal_host[2] = {nullptr, nullptr};
    size_t scal_host_cap[2] = {0, 0};
    cudaEvent_t scal_ev[2] = {nullptr, nullptr};
    int scal_slot = 0;

    // profiling
    bool profiling = false;
    std::vector<cudaEvent_t> ev;
    std::vector<int> ev_kind;

    ~ion_sim()
    {
        cudaSetDevice(device);
        void *ptrs[] = {psi,  io_stage, h_diag, h_off, w,    aggP, aggQ,       toff,        toff_prev,   vec,     zvec,   zprev,
                        mask, rvec,     cl,     cl2,   cl_z, scal, state_rows, state_first, state_order, partial, ip_out, obs_out};
        for (void *p : ptrs)
            if (p) cudaFree(p);
        if (psi2) cudaFree(psi2);
        if (slab_partial) cudaFree(slab_partial);
        if (slab_ip) cudaFree(slab_ip);
        for (auto q : obs_psi)
            if (q) cudaFree(q);
        if (partial2) cudaFree(partial2);
        if (ip_out2) cudaFree(ip_out2);
        if (obs_counter) cudaFree(obs_counter);
        if (ev_fork) cudaEventDestroy(ev_fork);
        for (auto e : ev_done)
            if (e) cudaEventDestroy(e);
        if (side) cudaStreamDestroy(side);
        if (scal_chunk) cudaFree(scal_chunk);
        if (scal_phase) cudaFree(scal_phase);
        if (th) cudaFree(th);
        if (thd) cudaFree(thd);
        if (obs_chunk) cudaFree(obs_chunk);
        for (void *q : peer_ipc_base)
            if (q) cudaIpcCloseMemHandle(q);
        if (hflags) cudaFree(hflags);
        if (hf_sent) cudaFree(hf_sent);
        for (int k = 0; k < 2; ++k) {
            if (scal_host[k]) cudaFreeHost(scal_host[k]);
            if (scal_ev[k]) cudaEventDestroy(scal_ev[k]);
        }
        for (auto &g : graphs)
            if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
        for (auto e : ev) cudaEventDestroy(e);
        if (own_stream && stream) cudaStreamDestroy(stream);
    }
    void invalidate_graphs()
    {
        for (auto &g : graphs)
            if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
        graphs.clear();
    }
};

namespace {

// host vector [R] -> device, permuted to [M][T], zero padded; rows >= limit are zeroed
int upload_permuted(ion_sim *s, const double *src, int n_valid, double **dst, double **dst_prev)
{
    std::vector<double> tmp((size_t)s->Rp, 0.0), prev((size_t)s->T, 0.0);
    for (int i = 0; i < n_valid; ++i) tmp[(size_t)(i % s->M) * s->T + (i / s->M)] = src[i];
    if (dst_prev)
        for (int t = 1; t < s->T; ++t) {
            int i = t * s->M - 1;
            prev[t] = i < n_valid ? src[i] : 0.0;
        }
    if (int rc = dev_alloc(dst, (size_t)s->Rp)) return rc;
    CUDA_TRY(cudaMemcpyAsync(*dst, tmp.data(), tmp.size() * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    if (dst_prev) {
        if (int rc = dev_alloc(dst_prev, (size_t)s->T)) return rc;
        CUDA_TRY(cudaMemcpyAsync(*dst_prev, prev.data(), prev.size() * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return ION_OK;
}

int upload_plain(ion_sim *s, const double *src, size_t n, double **dst)
{
    if (int rc = dev_alloc(dst, n)) return rc;
    if (n) {
        CUDA_TRY(cudaMemcpyAsync(*dst, src, n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
    }
    return ION_OK;
}

void prof_begin(ion_sim *s, int kind)
{
    if (!s->profiling) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, s->stream);
    s->ev.push_back(e);
    s->ev_kind.push_back(kind);
}
void prof_end(ion_sim *s)
{
    if (!s->profiling) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, s->stream);
    s->ev.push_back(e);
    s->ev_kind.push_back(-1);
}

// scan scratch + r-pair exchange; the Crank-Nicolson pair programs also stage the LU factors of both channels (layout 2)
size_t unit_smem_bytes(const ion_sim *s, int prog = -1)
{
    size_t n = (256 + 4 * (size_t)s->Tc) * sizeof(cplx);
    const bool cn_pair = (prog == ion::PROG_ROT_CN_ROT || prog == ion::PROG_H2_CN_H2 || prog == ion::PROG_LEN_STEP || prog == ion::PROG_LEN_STEP_OBS || prog == ion::PROG_LEN_STEP_HALO || prog < 0);
    if (cn_pair && s->M == 4 && s->tmax <= 512) n += (8 * (size_t)s->Tc + 1) * sizeof(cplx) + (9 * (size_t)(s->Tc / 2) + 1) * sizeof(double) + 16;  // + pad, mbarrier
    return n;
}

template <int PROG>
int launch_unit_prog(ion_sim *s, const ion::UnitParams &p, dim3 grid)
{
    const size_t smem = unit_smem_bytes(s, PROG);
    const dim3 block(PROG == ion::PROG_ROT ? s->T_seg : s->Tc);
    // programmatic dependent launch: the kernel's psi-independent prologue overlaps the previous kernel's tail
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = s->use_pdl ? 1 : 0;
#define ION_LAUNCH(MM, TMAX, SEG)                                                                                   \
    do {                                                                                                            \
        auto kern = ion::k_unit<MM, PROG, TMAX, SEG>;                                                               \
        CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, p));                                                                \
    } while (0)
    if (s->S > 1 && s->M == 8) {  // long LineMesh channels (Crank-Nicolson program): eight rows per thread, 256-thread CTAs
        if constexpr (PROG == ion::PROG_LINE_CN) ION_LAUNCH(8, 256, true);
        else return fail(ION_ESTATE, "internal: eight rows per thread in r-segments is built for the LineMesh Crank-Nicolson program only");
    } else if (s->S > 1) ION_LAUNCH(4, 512, true);  // r-segments: T_seg + 2H <= 512 threads
    else if (s->M == 8) ION_LAUNCH(8, 256, false);  // eight rows per thread, 256-thread CTAs (r_points <= 2048)
    else if (s->tmax == 256) ION_LAUNCH(4, 256, false);
    else if (s->tmax == 512) ION_LAUNCH(4, 512, false);
    else ION_LAUNCH(4, 1024, false);
#undef ION_LAUNCH
    CUDA_TRY(cudaGetLastError());
    return ION_OK;
}

template <int PROG>
int set_unit_smem_attr()
{
    CUDA_TRY(cudaFuncSetAttribute(ion::k_unit<4, PROG, 1024, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    if (PROG == ion::PROG_ROT_CN_ROT || PROG == ion::PROG_H2_CN_H2 || PROG == ion::PROG_LEN_STEP || PROG == ion::PROG_LEN_STEP_OBS || PROG == ion::PROG_LEN_STEP_HALO) {
        CUDA_TRY(cudaFuncSetAttribute(ion::k_unit<4, PROG, 512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
        CUDA_TRY(cudaFuncSetAttribute(ion::k_unit<4, PROG, 256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
        CUDA_TRY(cudaFuncSetAttribute(ion::k_unit<4, PROG, 512, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    }
    return ION_OK;
}
// kernels launched with more than 48 KB of dynamic shared memory (T > 704) need the opt-in; done once at creation,
// never inside a stream capture
int prepare_kernels(ion_sim *s)
{
    if (unit_smem_bytes(s) <= 48 * 1024) return ION_OK;
    int rc;
    if ((rc = set_unit_smem_attr<ion::PROG_ROT>()) || (rc = set_unit_smem_attr<ion::PROG_ROT_CN_ROT>()) ||
        (rc = set_unit_smem_attr<ion::PROG_H2>()) || (rc = set_unit_smem_attr<ion::PROG_H2_CN_H2>()) ||
        (rc = set_unit_smem_attr<ion::PROG_CN>()) || (rc = set_unit_smem_attr<ion::PROG_LINE_SO_LEN>()) ||
        (rc = set_unit_smem_attr<ion::PROG_LINE_SO_VEL>()) || (rc = set_unit_smem_attr<ion::PROG_LINE_CN>()) ||
        (rc = set_unit_smem_attr<ion::PROG_LEN_STEP>()) || (rc = set_unit_smem_attr<ion::PROG_LEN_STEP_OBS>()) ||
        (rc = set_unit_smem_attr<ion::PROG_LEN_STEP_HALO>()))
        return rc;
    return ION_OK;
}

ion::UnitParams base_params(ion_sim *s)
{
    ion::UnitParams p;
    std::memset(&p, 0, sizeof(p));
    p.psi = s->psi;
    p.psi_out = s->psi;
    p.w = s->w;
    p.aggP = s->aggP;
    p.aggQ = s->aggQ;
    p.th = s->th;
    p.toff = s->toff;
    p.toff_prev = s->toff_prev;
    p.vec = s->vec;
    p.zvec = s->zvec;
    p.zprev = s->zprev;
    p.mask = s->mask;
    p.cl = s->cl;
    p.cl2 = s->cl2;
    p.L = s->L;
    p.T = s->T;
    p.S = s->S;
    p.T_seg = s->T_seg;
    p.H = s->H;
    p.l_begin = s->l_begin;
    p.short_scan = s->short_scan;
    p.obs_rvec = s->rvec;
    p.obs_state_rows = s->state_rows;
    p.obs_state_first = s->state_first;
    p.obs_state_order = s->state_order;
    p.obs_partial = s->partial;
    p.obs_ip = s->ip_out;
    p.obs_ipm = s->ipm;
    p.vec_dv = s->vec_dv;
    p.unit0 = 0;
    p.unit_stride = 1;
    return p;
}

int launch_len_ens(ion_sim *s, const ion::UnitParams &p);
int launch_observe_finish(ion_sim *s, uint32_t what, double *dev_out);
int launch_observe(ion_sim *s, uint32_t what, double *dev_out);

// a subset of the units of a launch: units sub_unit0 + k * sub_stride, k < sub_count (sub_count == 0: all units).  do_swap = false:
// an out-of-place kernel leaves the buffer swap to the launch that covers the remaining units
int launch_unit(ion_sim *s, int prog, int parity, int flags, const double *sa, const double *sb, uint32_t obs_what = 0, int sub_unit0 = 0, int sub_stride = 1,
                int sub_count = 0, bool do_swap = true, int halo_consume = 0)
{
    ion::UnitParams p = base_params(s);
    p.parity = parity;
    p.flags = flags;
    p.obs_what = obs_what;
    p.obs_n_radii = (obs_what & ION_OBS_NORM_WITHIN) ? s->n_radii : 0;
    p.obs_n_states = (obs_what & ION_OBS_INNER_PRODUCTS) ? s->n_states : 0;
    for (int q = 0; q < p.obs_n_radii; ++q) p.obs_radii[q] = s->radii[q];
    if ((flags & ion::F_MASK) && !s->mask) p.flags &= ~ion::F_MASK;
    p.scal_a = sa;
    p.scal_b = sb;
    int units;
    int kind;
    switch (prog) {
        case ion::PROG_CN:
        case ion::PROG_LINE_SO_LEN:
        case ion::PROG_LINE_SO_VEL:
        case ion::PROG_LINE_CN: units = s->L; break;
        default: units = ion::num_units(s->L, s->l_begin, parity);
    }
    if (sub_count > 0) {
        p.unit0 = sub_unit0;
        p.unit_stride = sub_stride;
        units = sub_count;
    }
    dim3 grid(units * s->S, s->batch);
    if (prog == ion::PROG_ROT) p.H = 0;  // point-wise in r: interior threads only
    // r-segments: a kernel that reads halo rows (every program but the point-wise rotation) must not run in place
    // the ADI solve goes back to the buffer the l-pass read from, so that a step ends where it began
    const bool seg_oop = (s->S > 1 && prog != ion::PROG_ROT && prog != ion::PROG_LEN_STEP && prog != ion::PROG_LEN_STEP_OBS && prog != ion::PROG_LEN_STEP_HALO) || (prog == ion::PROG_CN && (flags & ion::F_SOLVE_ONLY));
    if (seg_oop) {
        if (!s->psi2) return fail(ION_ESTATE, "internal: second wavefunction buffer missing for a segmented kernel");
        p.psi_out = s->psi2;
    }
    int rc = ION_OK;
    switch (prog) {
        case ion::PROG_ROT:
            kind = KK_ROT;
            prof_begin(s, kind);
            rc = launch_unit_prog<ion::PROG_ROT>(s, p, grid);
            break;
        case ion::PROG_ROT_CN_ROT:
            kind = KK_ROT_CN_ROT;
            prof_begin(s, kind);
            rc = launch_unit_prog<ion::PROG_ROT_CN_ROT>(s, p, grid);
            break;
        case ion::PROG_H2:
            kind = KK_H2;
            prof_begin(s, kind);
            rc = launch_unit_prog<ion::PROG_H2>(s, p, grid);
            break;
        case ion::PROG_H2_CN_H2:
            kind = KK_H2_CN_H2;
            prof_begin(s, kind);
            rc = launch_unit_prog<ion::PROG_H2_CN_H2>(s, p, grid);
            break;
        case ion::PROG_CN:
            kind = KK_CN;
            prof_begin(s, kind);
            rc = launch_unit_prog<ion::PROG_CN>(s, p, grid);
            break;
        case ion::PROG_LINE_SO_LEN:
            kind = KK_LINE_SO_LEN;
            prof_begin(s, kind);
            rc = launch_unit_prog<ion::PROG_LINE_SO_LEN>(s, p, grid);
            break;
        case ion::PROG_LINE_SO_VEL:
            kind = KK_LINE_SO_VEL;
            prof_begin(s, kind);
            rc = launch_unit_prog<ion::PROG_LINE_SO_VEL>(s, p, grid);
            break;
        case ion::PROG_LINE_CN:
            kind = KK_LINE_CN;
            prof_begin(s, kind);
            rc = launch_unit_prog<ion::PROG_LINE_CN>(s, p, grid);
            break;
        case ion::PROG_LEN_STEP:
            kind = KK_LEN_STEP;
            p.psi_out = s->psi2;
            if (s->ens_state == 1) {
                rc = launch_len_ens(s, p);
            } else {
                prof_begin(s, kind);
                rc = launch_unit_prog<ion::PROG_LEN_STEP>(s, p, grid);
            }
            if (do_swap) std::swap(s->psi, s->psi2);
            break;
        case ion::PROG_LEN_STEP_HALO: {  // linked l-block shard cut at odd channels: the folded step with the halo exchange fused in
            kind = KK_LEN_STEP;
            p.psi_out = s->psi2;
            const size_t chan = (size_t)s->Rp;
            cplx *fstage = reinterpret_cast<cplx *>(s->hflags + ion::HF_COUNT) + 4 * chan;  // [2 sides][2 slots][Rp], behind the exchange kernel's slots
            p.hf_flags = s->hflags;
            for (int side = 0; side < 2; ++side) {
                p.hf_peer_flags[side] = s->peer_flags[side];
                p.hf_peer_fstage[side] = s->peer_fstage[side];
                p.hf_my_fstage[side] = fstage + (size_t)side * 2 * chan;
            }
            p.hf_sent = s->hf_sent;
            p.hf_spin_limit = 20000000000ll;  // ~10 s of SM clock
            p.hf_unit[0] = s->g_lo ? p.unit0 : -1;
            p.hf_unit[1] = s->g_hi ? p.unit0 + (units - 1) * p.unit_stride : -1;
            p.hf_consume = halo_consume;
            p.n_units = units;
            p.boundary_first = units >= 2 ? 1 : 0;
            prof_begin(s, kind);
            rc = launch_unit_prog<ion::PROG_LEN_STEP_HALO>(s, p, grid);
            if (do_swap) std::swap(s->psi, s->psi2);
            break;
        }
        case ion::PROG_LEN_STEP_OBS:  // never the persistent ensemble kernel: the observed step runs one CTA per (unit, member)
            kind = KK_LEN_STEP;
            p.psi_out = s->psi2;
            prof_begin(s, kind);
            rc = launch_unit_prog<ion::PROG_LEN_STEP_OBS>(s, p, grid);
            std::swap(s->psi, s->psi2);
            break;
        default: return fail(ION_EINVAL, "unknown unit program");
    }
    prof_end(s);
    s->launch_count++;
    if (seg_oop && rc == ION_OK && do_swap) std::swap(s->psi, s->psi2);
    return rc;
}

int launch_sweep_flat(ion_sim *s, int parity, int flags, const double *sa)
{
    if (s->L < 2) return ION_OK;
    ion::SweepParams p;
    std::memset(&p, 0, sizeof(p));
    p.psi = s->psi;
    p.vec = s->vec;
    p.cl = s->cl;
    p.scal_a = sa;
    p.L = s->L;
    p.L_total = s->L_total;
    p.l_begin = s->l_begin;
    p.T = s->T;
    p.M = s->M;
    p.R = s->R;
    p.parity = parity;
    p.flags = flags;
    dim3 block(128), grid((s->Rp + 127) / 128, s->L - 1, s->batch);
    prof_begin(s, KK_SWEEP_FLAT);
    ion::k_sweep_flat<<<grid, block, 0, s->stream>>>(p);
    prof_end(s);
    s->launch_count++;
    CUDA_TRY(cudaGetLastError());
    return ION_OK;
}

int launch_mask(ion_sim *s)
{
    if (!s->mask) return ION_OK;
    long long n = (long long)s->batch * s->L;
    dim3 block(128), grid((s->Rp + 127) / 128, (unsigned)std::min<long long>(n, 4096));
    prof_begin(s, KK_MASK);
    ion::k_mask<<<grid, block, 0, s->stream>>>(s->psi, s->mask, s->Rp, n);
    prof_end(s);
    s->launch_count++;
    CUDA_TRY(cudaGetLastError());
    return ION_OK;
}

// (re)build the LU factors of (1 + i tau H0) when tau changed
int ensure_factor(ion_sim *s, double tau)
{
    if (s->factored && std::fabs(tau - s->factored_tau) <= 1e-9 * std::fabs(tau)) return ION_OK;
    if (!s->have_h) return fail(ION_ESTATE, "ion_sim_set_hamiltonian must be called before stepping");
    s->invalidate_graphs();  // factor buffers are (re)allocated below
    // toff (host-side, tiny)
    std::vector<double> off(s->h_off_host);
    for (auto &v : off) v *= tau;
    if (int rc = upload_permuted(s, off.data(), s->R - 1, &s->toff, &s->toff_prev)) return rc;
    if (int rc = dev_alloc(&s->w, (size_t)s->L * s->Rp)) return rc;
    if (int rc = dev_alloc(&s->aggP, (size_t)s->L * s->T)) return rc;
    if (int rc = dev_alloc(&s->aggQ, (size_t)s->L * s->T)) return rc;
    ion::k_factor<<<(s->L + 31) / 32, 32, 0, s->stream>>>(s->h_diag, s->h_off, tau, s->L, s->R, s->M, s->T, s->w);
    CUDA_TRY(cudaGetLastError());
    dim3 g2((s->T + 127) / 128, s->L);
    ion::k_aggregates<<<g2, 128, 0, s->stream>>>(s->w, s->toff, s->L, s->M, s->T, s->aggP, s->aggQ);
    CUDA_TRY(cudaGetLastError());
    {   // is the cross-warp inflow of the scans short-ranged?  (product of multipliers over any warp < 1e-30)
        const int nw = s->T / 32, n = s->L * nw;
        double *d_bound = nullptr;
        if (int rc = dev_alloc(&d_bound, (size_t)n)) return rc;
        ion::k_scan_bound<<<(n + 127) / 128, 128, 0, s->stream>>>(s->aggP, s->aggQ, s->L, s->T, 32, d_bound);
        std::vector<double> hb((size_t)n);
        cudaError_t e = cudaMemcpyAsync(hb.data(), d_bound, hb.size() * sizeof(double), cudaMemcpyDeviceToHost, s->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
        cudaFree(d_bound);
        if (e != cudaSuccess) return fail(ION_ECUDA, std::string("k_scan_bound: ") + cudaGetErrorString(e));
        double mx = -1e300;
        for (int l = 0; l < s->L; ++l)
            for (int w = 0; w < nw; ++w) {
                // warp 0 in the forward direction / the last warp backwards have no inflow; their bound is irrelevant but harmless
                mx = std::max(mx, hb[(size_t)l * nw + w]);
            }
        // reach = number of whole warps over which the product of multipliers falls below 1e-30
        int reach = 0;
        if (nw > 1 && mx < 0.0) {
            reach = (int)std::ceil(std::log(1e30) / (-mx));
            if (reach > 2) reach = 0;  // slow decay: full block-wide scan
        }
        if (const char *env = std::getenv("ION_FULL_SCAN"))
            if (env[0] == '1') reach = 0;
        s->short_scan = reach;
        if (s->program == ION_LINE_LEN_CN) {
            if (nw > 1 && reach == 0)
                return fail(ION_ENOTSUP,
                            "LineMesh Crank-Nicolson rebuilds its pivots per warp and needs the LU multipliers to decay below 1e-30 "
                            "over 256 rows; this time step is too large for the mesh spacing");
            {   // the bound above is field-free; with the field the diagonal carries tau E w_z as well, and where that cancels the kinetic
                // diagonal the multipliers decay slowest: |lambda| = (sqrt(1 + 4 a^2) - 1) / (2 a), a = tau |h_off|.  The truncations of this
                // program (128-row pivot memory, 256-row inflow / segment halos) must stay below double precision for ANY field.
                double a = 0.0;
                for (double v : s->h_off_host) a = std::max(a, std::fabs(tau * v));
                const double lam = a > 0.0 ? (std::sqrt(1.0 + 4.0 * a * a) - 1.0) / (2.0 * a) : 0.0;
                if (nw > 1 && (std::pow(lam, 256.0) > 1e-16 || std::pow(lam * lam, 128.0) > 1e-16))
                    return fail(ION_ENOTSUP,
                                "LineMesh Crank-Nicolson: tau * h_off = " + std::to_string(a) + " is too large for the truncated scans of this program to "
                                "stay exact when the field's potential cancels the kinetic diagonal (reduce time_step or increase the spacing)");
            }
            if (int rc = dev_alloc(&s->th, (size_t)s->Rp)) return rc;
            ion::k_make_th<<<(s->Rp + 127) / 128, 128, 0, s->stream>>>(s->h_diag, tau, s->R, s->M, s->T, s->th);
            CUDA_TRY(cudaGetLastError());
        }
        if (s->program == ION_SH_LEN_ADI) {
            if (int rc = dev_alloc(&s->thd, (size_t)s->L * s->Rp)) return rc;
            dim3 g3((s->Rp + 127) / 128, s->L);
            ion::k_make_thd<<<g3, 128, 0, s->stream>>>(s->h_diag, tau, s->R, s->M, s->T, s->thd);
            CUDA_TRY(cudaGetLastError());
        }
        if (s->S > 1) {
            // r-segments: the halo must cover the reach; programs with r-pair bricks (velocity gauge) get one more warp of
            // margin per unit of reach for the brick edge effects
            const bool bricks = (s->program == ION_SH_VEL_SO || s->program == ION_LINE_VEL_SO);
            int H = (bricks ? 64 : 32) * reach;
            if (s->program == ION_SH_LEN_SO && reach == 1 && s->T_seg % 32 == 0) {
                // half-warp halos: the product of the multipliers over any aligned 16 threads (64 rows) is below 1e-18, i.e. what a
                // segment ignores of its neighbour is two orders below the rounding of the values it would multiply
                const char *env = std::getenv("ION_HALO16");
                if (!(env && env[0] == '0')) {
                    const int ng = s->T / 16, n = s->L * ng;
                    double *d_bound = nullptr;
                    if (int rc = dev_alloc(&d_bound, (size_t)n)) return rc;
                    ion::k_scan_bound<<<(n + 127) / 128, 128, 0, s->stream>>>(s->aggP, s->aggQ, s->L, s->T, 16, d_bound);
                    std::vector<double> hb((size_t)n);
                    cudaError_t e = cudaMemcpyAsync(hb.data(), d_bound, hb.size() * sizeof(double), cudaMemcpyDeviceToHost, s->stream);
                    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
                    cudaFree(d_bound);
                    if (e != cudaSuccess) return fail(ION_ECUDA, std::string("k_scan_bound: ") + cudaGetErrorString(e));
                    double mx16 = -1e300;
                    for (double v : hb) mx16 = std::max(mx16, v);
                    if (mx16 < std::log(1e-18)) H = 16;
                    if (std::getenv("ION_DEBUG")) std::fprintf(stderr, "[ion] r-segments: multipliers over 16 threads <= %.3g, halo %d threads\n", std::exp(mx16), H);
                }
            }
            if (reach == 0 || s->T_seg + 2 * H > (s->M == 8 ? 256 : 512))
                return fail(ION_ENOTSUP,
                            "r_points > 4096 needs the Crank-Nicolson LU multipliers to decay below 1e-30 within the segment halo; this "
                            "time step is too large for the radial spacing (reduce time_step or increase the spacing)");
            s->H = H;
            s->Tc = s->T_seg + 2 * H;
        }
    }
    s->factored = true;
    s->factored_tau = tau;
    return ION_OK;
}

bool fast_l_path(const ion_sim *s) { return (s->L_total % 2) == 0; }

// ---------------------------------------------------------------------------------------------
// velocity-gauge inter-solve kernel (slab.cuh): one out-of-place pass replaces the five pair-local passes between two
// Crank-Nicolson solves.  Decided once per handle: split-operator velocity gauge, even l_bound, unsharded, M = 4, S = 1.
// ---------------------------------------------------------------------------------------------
int ensure_second_buffer(ion_sim *s)
{
    if (s->psi2) return ION_OK;
    const size_t n = (size_t)s->batch * s->L * s->Rp;
    if (int rc = dev_alloc(&s->psi2, n)) return rc;
    CUDA_TRY(cudaMemsetAsync(s->psi2, 0, n * sizeof(cplx), s->stream));  // the padding rows are never written by the slab kernel
    s->psi_home = s->psi;
    return ION_OK;
}

// length gauge: the even sweep (tail of step n-1, mask, head of step n) is folded into the odd-pair Crank-Nicolson kernel,
// which reads the even-pair partners of its two channels read-only and therefore works out of place: one pass per step.
int len_fold_prepare(ion_sim *s)
{
    if (s->len_fold_state != 0) return ION_OK;
    s->len_fold_state = -1;
    if (!s->use_len_fold || s->program != ION_SH_LEN_SO || !fast_l_path(s)) return ION_OK;
    if (s->L < 2) return ION_OK;
    // l-block shards: cut at odd channels and linked by the engine's own exchange (the phase API stays on the single-sweep kernels)
    if (s->L_own != s->L_total && !(s->cut_parity == 1 && s->peers_attached)) return ION_OK;
    if (int rc = ensure_second_buffer(s)) return rc;
    s->len_fold_state = 1;
    // ensembles: persistent CTAs with a prefetch pipeline instead of one short-lived CTA per (pair, member)
    s->ens_state = -1;
    if (s->use_ens && s->M == 4 && s->S == 1 && s->T == ion::ENS_T && s->L >= 4 && !s->peers_attached) {
        const long long tasks = (long long)s->batch * (s->L / 2 - 1);
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, s->device));
        CUDA_TRY(cudaFuncSetAttribute(ion::k_len_ens<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ion::ens_smem_bytes(ion::ENS_T)));
        CUDA_TRY(cudaFuncSetAttribute(ion::k_len_ens<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ion::ens_smem_bytes(ion::ENS_T)));
        int per_sm = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ion::k_len_ens<false>, ion::ENS_T, ion::ens_smem_bytes(ion::ENS_T)));
        if (per_sm >= 1 && tasks >= 4LL * per_sm * prop.multiProcessorCount) {
            s->ens_ctas = per_sm * prop.multiProcessorCount;
            s->ens_state = 1;
        }
    }
    return ION_OK;
}

// the folded length-gauge step of an ensemble: persistent kernel over the channel pairs + k_unit over the two single channels
int launch_len_ens(ion_sim *s, const ion::UnitParams &p)
{
    const int n_pairs = s->L / 2 - 1;
    const long long n_items = (long long)n_pairs * ((s->batch + ion::ENS_MB - 1) / ion::ENS_MB);
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)std::min<long long>(n_items, s->ens_ctas));
    cfg.blockDim = dim3(s->T);
    cfg.dynamicSmemBytes = ion::ens_smem_bytes(s->T);
    cfg.stream = s->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = s->use_pdl ? 1 : 0;
    prof_begin(s, KK_LEN_ENS);
    {
        // psi prefetch by the TMA engine (cp.async.bulk + mbarrier): 1095 -> 998 us per step on configs[3] (45.7 -> 50.1 % of the roofline);
        // ION_ENS_BULK=0 keeps the 16 cp.async per thread for A/B timing
        const char *env = std::getenv("ION_ENS_BULK");
        if (env && env[0] == '0') CUDA_TRY(cudaLaunchKernelEx(&cfg, ion::k_len_ens<false>, p, n_pairs, s->batch));
        else CUDA_TRY(cudaLaunchKernelEx(&cfg, ion::k_len_ens<true>, p, n_pairs, s->batch));
    }
    prof_end(s);
    s->launch_count++;
    ion::UnitParams q = p;  // l = 0 and l = L - 1: units 0 and L/2 of the odd sweep
    q.unit0 = 0;
    q.unit_stride = s->L / 2;
    prof_begin(s, KK_LEN_STEP);
    return launch_unit_prog<ion::PROG_LEN_STEP>(s, q, dim3(2, s->batch));
}

int slab_prepare(ion_sim *s)
{
    if (s->slab_state != 0) return ION_OK;
    s->slab_state = -1;
    if (!s->use_slab || s->program != ION_SH_VEL_SO || !fast_l_path(s)) return ION_OK;
    if (s->L_own != s->L_total || s->M != 4 || s->S != 1 || s->L < 2) return ION_OK;
    int G = 8;
    if (const char *env = std::getenv("ION_SLAB_G")) {
        const int v = std::atoi(env);
        if (v == 4 || v == 8 || v == 16 || v == 32) G = v;
    }
    const int nt_cap = 512;  // one 512-thread CTA per SM at 128 registers (two co-resident 288-thread CTAs at 96 registers spill: measured 32.3 vs 27.6 us per VEL step)
    const int nQ = (s->L + 3) / 4;
    int chunks = 1, Qc = nQ, loaded = nQ;
    for (;; ++chunks) {
        Qc = (nQ + chunks - 1) / chunks;
        loaded = Qc + (chunks == 1 ? 0 : (chunks == 2 ? 1 : 2));
        if (G * loaded <= nt_cap || Qc == 1) break;
    }
    if (G * loaded > nt_cap) return ION_OK;
    chunks = (nQ + Qc - 1) / Qc;
    const int W = 4 * G - 4;
    s->slab_G = G;
    s->slab_Qc = Qc;
    s->slab_nQ = nQ;
    s->slab_chunks = chunks;
    s->slab_slabs = s->R / W + 1;
    s->slab_threads = (G * loaded + 31) / 32 * 32;
    if (int rc = ensure_second_buffer(s)) return rc;
    CUDA_TRY(cudaFuncSetAttribute(ion::k_slab<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * 512 * (int)sizeof(cplx)));
    CUDA_TRY(cudaFuncSetAttribute(ion::k_slab<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * 512 * (int)sizeof(cplx)));
    CUDA_TRY(cudaFuncSetAttribute(ion::k_slab<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * 512 * (int)sizeof(cplx)));
    s->slab_state = 1;
    return ION_OK;
}

// obs_what != 0: the state after this step's mask is observed INSIDE the kernel (slab.cuh: SlabObs) and the record goes to obs_dst
int launch_slab(ion_sim *s, const double *sa, const double *sb, uint32_t obs_what, double *obs_dst)
{
    ion::SlabParams p;
    std::memset(&p, 0, sizeof(p));
    p.psi_in = s->psi;
    p.psi_out = s->psi2;
    p.vec = s->vec;
    p.zvec = s->zvec;
    p.mask = s->mask;
    p.cl = s->cl;
    p.cl2 = s->cl2;
    p.scal_a = sa;
    p.scal_b = sb;
    p.L = s->L;
    p.T = s->T;
    p.R = s->R;
    p.G = s->slab_G;
    p.n_slabs = s->slab_slabs;
    p.Qc = s->slab_Qc;
    p.nQ = s->slab_nQ;
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(s->slab_slabs * s->slab_chunks, s->batch);
    cfg.blockDim = dim3(s->slab_threads);
    cfg.dynamicSmemBytes = 12 * (size_t)s->slab_threads * sizeof(cplx);
    cfg.stream = s->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = s->use_pdl ? 1 : 0;
    ion::SlabObs o;
    std::memset(&o, 0, sizeof(o));
    prof_begin(s, KK_SLAB);
    if (!obs_what) {
        CUDA_TRY(cudaLaunchKernelEx(&cfg, ion::k_slab<false>, p, o));
    } else {
        if (!s->slab_partial || !s->slab_ip) return fail(ION_ESTATE, "internal: fused-observation buffers missing");
        o.rvec = s->rvec;
        o.state_rows = s->state_rows;
        o.state_first = s->state_first;
        o.state_order = s->state_order;
        const int k = s->obs_parity;
        o.partial = s->slab_partial + (size_t)k * s->slab_partial_half;
        o.ip = s->slab_ip + (size_t)k * s->slab_ip_half;
        if (s->side_pending[k]) {  // the assembly that last read this half (two observations ago) must be done
            CUDA_TRY(cudaStreamWaitEvent(s->stream, s->ev_done[k], 0));
            s->side_pending[k] = false;
        }
        o.n_radii = (obs_what & ION_OBS_NORM_WITHIN) ? s->n_radii : 0;
        o.n_states = (obs_what & ION_OBS_INNER_PRODUCTS) ? s->n_states : 0;
        for (int q = 0; q < o.n_radii; ++q) o.radii[q] = s->radii[q];
        o.what = obs_what;
        if (s->obs_psi[0] && s->obs_psi[1]) {
            o.psi_n = s->obs_psi[k];
            CUDA_TRY(cudaLaunchKernelEx(&cfg, ion::k_slab<true, true>, p, o));
        } else {
            CUDA_TRY(cudaLaunchKernelEx(&cfg, ion::k_slab<true>, p, o));
        }
    }
    prof_end(s);
    s->launch_count++;
    std::swap(s->psi, s->psi2);
    if (obs_what && s->obs_psi[0] && s->obs_psi[1]) {
        // side branch: k_observe + k_observe_finish on the stored state, while the main stream goes on with the next step
        const int k = s->obs_parity;
        s->obs_parity ^= 1;
        CUDA_TRY(cudaEventRecord(s->ev_fork, s->stream));
        CUDA_TRY(cudaStreamWaitEvent(s->side, s->ev_fork, 0));
        cudaStream_t main_stream = s->stream;
        cplx *main_psi = s->psi;
        const int g_lo = s->g_lo;
        s->stream = s->side;
        s->psi = s->obs_psi[k];
        int rc = launch_observe(s, obs_what, obs_dst);
        s->stream = main_stream;
        s->psi = main_psi;
        (void)g_lo;
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(s->ev_done[k], s->side));
        s->side_pending[k] = true;
        return ION_OK;
    }
    if (obs_what) {
        // side branch: assemble the record while the main stream goes on with the next step
        const int k = s->obs_parity;
        s->obs_parity ^= 1;
        CUDA_TRY(cudaEventRecord(s->ev_fork, s->stream));
        CUDA_TRY(cudaStreamWaitEvent(s->side, s->ev_fork, 0));
        ion::k_slab_obs_assemble<<<dim3(s->L + (2 * o.n_states + 63) / 64, s->batch), 128, 0, s->side>>>(
            o.partial, o.ip, s->partial, s->ip_out, s->obs_counter, obs_dst, s->slab_slabs, s->L, o.n_states, o.n_radii, (unsigned)obs_what, s->ipm,
            (long long)ion_sim_observation_size(s, obs_what));
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(s->ev_done[k], s->side));
        s->side_pending[k] = true;
        s->launch_count++;
    }
    return ION_OK;
}

// ION_SH_LEN_ADI: explicit r half-step, implicit + explicit l half-steps in one out-of-place pass (adi.cuh)
int launch_adi_l(ion_sim *s, const double *sa)
{
    if (!s->psi2 || !s->thd) return fail(ION_ESTATE, "internal: ADI buffers missing");
    ion::AdiParams p;
    std::memset(&p, 0, sizeof(p));
    p.psi = s->psi;
    p.out = s->psi2;
    p.thd = s->thd;
    p.toff = s->toff;
    p.toff_prev = s->toff_prev;
    p.vec = s->vec;
    p.cl = s->cl;
    p.scal = sa;
    p.L = s->L;
    p.T = s->T;
    p.M = s->M;
    p.NC = (s->L + ion::ADI_CL - 1) / ion::ADI_CL;
    int pw = 32;
    while (pw > 1 && pw * p.NC > ion::ADI_MAX_THREADS) pw /= 2;
    // two co-resident CTAs of <= 256 threads overlap each other's load / scan / store phases (2000 x 500: 35.8 -> 32.3 us per step with
    // 4 instead of 8 positions per CTA), as long as a CTA still covers 64 contiguous bytes of every channel
    if (pw >= 8 && (pw / 2) * p.NC <= 256) pw /= 2;
    if (const char *env = std::getenv("ION_ADI_PW")) {  // A/B switch: fewer positions per CTA (two co-resident CTAs per SM)
        const int v = std::atoi(env);
        if ((v == 1 || v == 2 || v == 4 || v == 8 || v == 16) && v <= pw) pw = v;
    }
    p.PW = pw;
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(s->Rp / pw, s->batch);
    cfg.blockDim = dim3((pw * p.NC + 31) / 32 * 32);  // whole warps: the prefixes shuffle
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = s->use_pdl ? 1 : 0;
    prof_begin(s, KK_ADI_L);
    CUDA_TRY(cudaLaunchKernelEx(&cfg, ion::k_adi_l, p));
    prof_end(s);
    s->launch_count++;
    std::swap(s->psi, s->psi2);
    return ION_OK;
}

// the closing radial solve of an ADI step with eight rows per thread (adi.cuh: k_adi_r): channels of at most 2048 points in the
// M = 4 layout, one CTA of T/2 threads per channel; everything else keeps k_unit<PROG_CN>
bool adi_r_ok(const ion_sim *s)
{
    const char *env = std::getenv("ION_NO_ADI_R");
    return !(env && env[0] == '1') && s->M == 4 && s->S == 1 && s->T <= 512 && s->T >= 64;
}

int launch_adi_r(ion_sim *s)
{
    if (!s->psi2) return fail(ION_ESTATE, "internal: ADI buffers missing");
    ion::AdiRParams p;
    std::memset(&p, 0, sizeof(p));
    p.psi = s->psi;
    p.out = s->psi2;
    p.w = s->w;
    p.aggP = s->aggP;
    p.aggQ = s->aggQ;
    p.toff = s->toff;
    p.toff_prev = s->toff_prev;
    p.mask = s->mask;
    p.L = s->L;
    p.T = s->T;
    p.short_scan = s->short_scan;
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(s->L, s->batch);
    cfg.blockDim = dim3((s->T / 2 + 31) / 32 * 32);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = s->use_pdl ? 1 : 0;
    prof_begin(s, KK_CN);
    CUDA_TRY(cudaLaunchKernelEx(&cfg, ion::k_adi_r, p));
    prof_end(s);
    s->launch_count++;
    std::swap(s->psi, s->psi2);
    return ION_OK;
}

// bring the current state back into the buffer the rest of the API (and every captured graph) starts from
int restore_home(ion_sim *s)
{
    if (!s->psi_home || s->psi == s->psi_home) return ION_OK;
    CUDA_TRY(cudaMemcpyAsync(s->psi2, s->psi, (size_t)s->batch * s->L * s->Rp * sizeof(cplx), cudaMemcpyDeviceToDevice, s->stream));
    std::swap(s->psi, s->psi2);
    s->launch_count++;
    return ION_OK;
}

// l-block shard with attached neighbours: deliver the boundary channels into the neighbours' ghosts (halo.cuh)
int launch_exchange(ion_sim *s)
{
    if (!s->peers_attached) return ION_OK;
    ion::HaloParams p;
    std::memset(&p, 0, sizeof(p));
    p.flags = s->hflags;
    const size_t chan = (size_t)s->Rp;
    cplx *stage = reinterpret_cast<cplx *>(s->hflags + ion::HF_COUNT);
    for (int side = 0; side < 2; ++side) {
        p.peer_flags[side] = s->peer_flags[side];
        p.peer_stage[side] = s->peer_stage[side];
        p.my_stage[side] = stage + (size_t)side * 2 * chan;
    }
    p.src[0] = s->psi + (size_t)s->g_lo * chan;
    p.src[1] = s->psi + (size_t)(s->g_lo + s->L_own - 1) * chan;
    p.ghost[0] = s->psi;
    p.ghost[1] = s->psi + (size_t)(s->g_lo + s->L_own) * chan;
    p.n = (long long)chan;
    p.spin_limit = 20000000000ll;  // ~10 s of SM clock
    prof_begin(s, KK_HALO);
    ion::k_halo_exchange<<<dim3(8, 2), 256, 0, s->stream>>>(p);
    prof_end(s);
    s->launch_count++;
    CUDA_TRY(cudaGetLastError());
    return ION_OK;
}

// A kernel of a linked l-block shard whose pairs straddle the cuts (parity != cut_parity), or the folded length-gauge step of a
// shard cut at odd channels (its first and last pair read a ghost channel as their even-pair partner).  Only the first and the
// last unit touch a ghost channel, so only they have to wait for the halo exchange: the exchange (NVLink latency + rendezvous with
// both neighbours) and those two units run on a side branch while the main stream does all the interior units -- the hand-shake
// is off the critical path as long as the neighbours are less than one interior kernel apart.  Works alike for plain launches and
// inside a stream capture (the event record / wait pairs become graph edges).  Used when every neighbour lives on another device
// (one shard per GPU): with several shards on ONE device a spinning exchange kernel at the head of a hardware work queue can hold
// back another shard's kernels that the driver mapped to the same queue, so those (tests, devices=[0, 0, ..]) keep one stream per shard.
bool fused_halo_ok(const ion_sim *s)
{
    // neighbours on this device (several shards of one process on one GPU: tests) keep the stand-alone exchange unless asked
    // (ION_FUSED_HALO=1): a CTA that spins for a neighbour's kernel shares the device with that kernel there
    const char *env = std::getenv("ION_FUSED_HALO");
    if (env && env[0] == '0') return false;
    if (s->neighbour_on_same_device && !(env && env[0] == '1')) return false;
    // the exchange lives in the layout-2 pair path of k_unit (CTAs of at most 512 threads; r-segments always are)
    return s->use_fused_halo && s->peers_attached && s->cut_parity == 1 && s->M == 4 && (s->S > 1 || s->tmax <= 512) && s->hf_sent && !s->profiling;
}

int launch_exchanged(ion_sim *s, int prog, int parity, int flags, const double *sa, const double *sb, bool consume = false)
{
    const bool folded = (prog == ion::PROG_LEN_STEP);
    if (!s->peers_attached || (!folded && parity == s->cut_parity)) return launch_unit(s, prog, parity, flags, sa, sb);
    // units of the launch: the folded step skips the ghost channels (single-channel units at either end; they belong to the neighbours)
    const int units_all = ion::num_units(s->L, s->l_begin, parity);
    const int u0 = folded ? s->g_lo : 0, u1 = units_all - 1 - (folded ? s->g_hi : 0);
    const int units = u1 - u0 + 1;
    if (folded && fused_halo_ok(s)) {
        // the exchange rides inside the step kernel: launch k stores its boundary channels into the neighbours' slots k & 1 and
        // reads its ghost partners from the slots the neighbours' launch k - 1 filled.  `consume` == false (the first folded step
        // after anything else): the ghost channels come from the stand-alone exchange, as before; the launch still sends.
        if (!consume)
            if (int rc = launch_exchange(s)) return rc;
        return launch_unit(s, ion::PROG_LEN_STEP_HALO, parity, flags, sa, sb, 0, u0, 1, units, true, consume ? 1 : 0);
    }
    const char *env = std::getenv("ION_SERIAL_EXCHANGE");
    if (!s->side || units < 3 || s->profiling || (env && env[0] == '1') || s->neighbour_on_same_device) {
        if (int rc = launch_exchange(s)) return rc;
        return launch_unit(s, prog, parity, flags, sa, sb, 0, u0, 1, units, true);
    }
    CUDA_TRY(cudaEventRecord(s->ev_fork, s->stream));
    CUDA_TRY(cudaStreamWaitEvent(s->side, s->ev_fork, 0));
    cudaStream_t main_stream = s->stream;
    s->stream = s->side;
    int rc = launch_exchange(s);
    if (rc == ION_OK) rc = launch_unit(s, prog, parity, flags, sa, sb, 0, u0, units - 1, 2, false);  // units u0 and u1
    s->stream = main_stream;
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(s->ev_done[0], s->side));
    if ((rc = launch_unit(s, prog, parity, flags, sa, sb, 0, u0 + 1, 1, units - 2, true))) return rc;               // the interior units
    CUDA_TRY(cudaStreamWaitEvent(s->stream, s->ev_done[0], 0));
    return ION_OK;
}

// what a step's tail does of the next step's head when the two are fused (no observation in between)
int fuse_level(const ion_sim *s) { return s->slab_state == 1 ? 2 : 1; }

// one step; `pre`: how much of this step's head was already done by the previous step's tail (0: nothing, 1: the leading
// even rotation, 2: velocity gauge with the inter-solve kernel -- everything up to the odd-pair Crank-Nicolson kernel);
// `fuse_next`: fold the head of the next step into this step's tail (to fuse_level()).
// A fused observation (north_star 4): `what` != 0 asks the kernels of this step to reduce an observed state on the fly and
// to leave the record at `dst`.  cur: the state after THIS step's mask (velocity gauge: inside k_slab); prev: the state after
// the PREVIOUS step's mask, whose trailing even rotation and mask were deferred into this step's kernel (length gauge).
struct ObsReq {
    uint32_t what = 0;
    double *dst = nullptr;
};

int enqueue_step(ion_sim *s, const double *sa, const double *sb_next, int pre, bool fuse_next, ObsReq prev = ObsReq(), ObsReq cur = ObsReq())
{
    using namespace ion;
    int rc = ION_OK;
    switch (s->program) {
        case ION_SH_LEN_SO:
            if (fast_l_path(s) && s->len_fold_state == 1) {
                // the previous step's deferred tail (its scalar is the row before sa), the mask and this step's head ride along
                if (prev.what && pre) {
                    // the record is assembled on a side branch (event fork / join, also inside the captured graphs): the next step's kernel
                    // depends on this step's kernel only; the partial-sum buffers alternate between two sets
                    const bool on_side = s->side && s->partial2 && s->ip_out2 && !s->profiling;
                    const int k = s->obs_parity;
                    if (on_side) {
                        if (s->side_pending[k]) {  // the assembly that last read this set (two observations ago) must be done
                            CUDA_TRY(cudaStreamWaitEvent(s->stream, s->ev_done[k], 0));
                            s->side_pending[k] = false;
                        }
                        if (k) std::swap(s->partial, s->partial2), std::swap(s->ip_out, s->ip_out2);
                    }
                    rc = launch_unit(s, PROG_LEN_STEP_OBS, 1, F_MASK, sa, sa - s->batch, prev.what);
                    if (rc == ION_OK && on_side) {
                        cudaStream_t main_stream = s->stream;
                        if (cudaEventRecord(s->ev_fork, main_stream) != cudaSuccess || cudaStreamWaitEvent(s->side, s->ev_fork, 0) != cudaSuccess)
                            rc = fail(ION_ECUDA, "fork of the observation branch failed");
                        if (rc == ION_OK) {
                            s->stream = s->side;
                            rc = launch_observe_finish(s, prev.what, prev.dst);
                            s->stream = main_stream;
                        }
                        if (rc == ION_OK && cudaEventRecord(s->ev_done[k], s->side) != cudaSuccess) rc = fail(ION_ECUDA, "cudaEventRecord failed");
                        if (rc == ION_OK) s->side_pending[k] = true, s->obs_parity ^= 1;
                    } else if (rc == ION_OK) {
                        prof_begin(s, KK_OBSERVE);
                        rc = launch_observe_finish(s, prev.what, prev.dst);
                        prof_end(s);
                    }
                    if (on_side && k) std::swap(s->partial, s->partial2), std::swap(s->ip_out, s->ip_out2);
                    if (rc) return rc;
                } else if ((rc = launch_exchanged(s, PROG_LEN_STEP, 1, pre ? F_MASK : 0, sa, pre ? sa - s->batch : nullptr, pre != 0))) return rc;
                return fuse_next ? ION_OK : launch_exchanged(s, PROG_ROT, 0, F_MASK, sa, nullptr);
            }
            if (fast_l_path(s)) {
                if (pre < 1 && (rc = launch_exchanged(s, PROG_ROT, 0, 0, sa, nullptr))) return rc;
                if ((rc = launch_exchanged(s, PROG_ROT_CN_ROT, 1, 0, sa, nullptr))) return rc;
                return launch_exchanged(s, PROG_ROT, 0, F_MASK, sa, fuse_next ? sb_next : nullptr);
            }
            if ((rc = launch_sweep_flat(s, 0, 0, sa))) return rc;
            if ((rc = launch_sweep_flat(s, 1, 0, sa))) return rc;
            if ((rc = launch_unit(s, PROG_CN, 0, 0, nullptr, nullptr))) return rc;
            if ((rc = launch_sweep_flat(s, 1, 0, sa))) return rc;
            if ((rc = launch_sweep_flat(s, 0, 0, sa))) return rc;
            return launch_mask(s);
        case ION_SH_VEL_SO: {
            const bool fast = fast_l_path(s);
            if (fast) {
                if (pre < 1 && (rc = launch_unit(s, PROG_ROT, 0, F_REAL_ROT, sa, nullptr))) return rc;
                if (pre < 2 && (rc = launch_exchanged(s, PROG_ROT, 1, F_REAL_ROT, sa, nullptr))) return rc;
            } else {
                if ((rc = launch_sweep_flat(s, 0, F_REAL_ROT, sa))) return rc;
                if ((rc = launch_sweep_flat(s, 1, F_REAL_ROT, sa))) return rc;
            }
            if (pre < 2 && (rc = launch_unit(s, PROG_H2, 0, 0, sa, nullptr))) return rc;   // ee, eo
            if ((rc = launch_exchanged(s, PROG_H2_CN_H2, 1, 0, sa, nullptr))) return rc;     // oe, oo, CN, oo, oe
            if (fast && fuse_next && s->slab_state == 1) return launch_slab(s, sa, sb_next, cur.what, cur.dst);  // eo ee h1_o h1_e mask | h1_e h1_o ee eo
            if ((rc = launch_unit(s, PROG_H2, 0, F_H2_REVERSE, sa, nullptr))) return rc;   // eo, ee
            if (fast) {
                if ((rc = launch_exchanged(s, PROG_ROT, 1, F_REAL_ROT, sa, nullptr))) return rc;
                return launch_unit(s, PROG_ROT, 0, F_REAL_ROT | F_MASK, sa, fuse_next ? sb_next : nullptr);
            }
            if ((rc = launch_sweep_flat(s, 1, F_REAL_ROT, sa))) return rc;
            if ((rc = launch_sweep_flat(s, 0, F_REAL_ROT, sa))) return rc;
            return launch_mask(s);
        }
        case ION_SH_LEN_ADI:
            // (1 - i tau H0)_r, (1 + i tau Hint)^-1_l, (1 - i tau Hint)_l | (1 + i tau H0)^-1_r, mask   (evolution_methods.py:49-77)
            if ((rc = launch_adi_l(s, sa))) return rc;
            if (adi_r_ok(s)) return launch_adi_r(s);
            return launch_unit(s, PROG_CN, 0, F_MASK | F_SOLVE_ONLY, nullptr, nullptr);
        case ION_LINE_LEN_SO: return launch_unit(s, PROG_LINE_SO_LEN, 0, F_MASK, sa, nullptr);
        case ION_LINE_VEL_SO: return launch_unit(s, PROG_LINE_SO_VEL, 0, F_MASK, sa, nullptr);
        case ION_LINE_LEN_CN: return launch_unit(s, PROG_LINE_CN, 0, F_MASK, sa, nullptr);
        default: return fail(ION_ENOTSUP, "program not supported by this build");
    }
}

bool program_fuses(const ion_sim *s)
{
    return (s->program == ION_SH_LEN_SO || s->program == ION_SH_VEL_SO) && fast_l_path(s);
}

int check_ready(ion_sim *s)
{
    if (!s->have_h) return fail(ION_ESTATE, "ion_sim_set_hamiltonian must be called before stepping");
    if (!s->have_coupling) return fail(ION_ESTATE, "the coupling vectors of this program have not been set");
    return ION_OK;
}

int upload_scalars(ion_sim *s, int64_t n_steps, const double *taus, const double *fields)
{
    // row 0 and row n_steps + 1 are zero: "the step before the first" / "the step after the last"; step n is row n + 1
    size_t need = (size_t)(n_steps + 2) * s->batch;
    if (need > s->scal_cap) {
        if (int rc = dev_alloc(&s->scal, need)) return rc;
        s->scal_cap = need;
    }
    // pinned staging, two slots: the copy is asynchronous and the host never waits for the stream here -- only for the
    // copy that last used the slot it is about to refill (two calls ago)
    const int k = s->scal_slot;
    s->scal_slot ^= 1;
    if (!s->scal_ev[k]) CUDA_TRY(cudaEventCreateWithFlags(&s->scal_ev[k], cudaEventDisableTiming));
    else CUDA_TRY(cudaEventSynchronize(s->scal_ev[k]));
    if (need > s->scal_host_cap[k]) {
        if (s->scal_host[k]) cudaFreeHost(s->scal_host[k]);
        s->scal_host[k] = nullptr;
        s->scal_host_cap[k] = 0;
        CUDA_TRY(cudaMallocHost((void **)&s->scal_host[k], need * sizeof(double)));
        s->scal_host_cap[k] = need;
    }
    double *h = s->scal_host[k];
    std::memset(h, 0, need * sizeof(double));
    for (int64_t n = 0; n < n_steps; ++n)
        for (int b = 0; b < s->batch; ++b) h[(size_t)(n + 1) * s->batch + b] = taus[n] * fields[(size_t)n * s->batch + b];
    CUDA_TRY(cudaMemcpyAsync(s->scal, h, need * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaEventRecord(s->scal_ev[k], s->stream));
    return ION_OK;
}

// after a synchronisation point: did a peer-memory halo exchange of this l-block shard time out?  A shard that missed a
// hand-shake has kept stepping with stale ghost channels, so the wavefunction of the handle is invalid from then on.
int check_abort(ion_sim *s)
{
    if (!s->peers_attached || !s->hflags) return ION_OK;
    unsigned long long flag = 0;
    CUDA_TRY(cudaMemcpy(&flag, s->hflags + ion::HF_ABORT, sizeof(flag), cudaMemcpyDeviceToHost));
    if (flag)
        return fail(ION_ECUDA, "l-block shard: a peer-memory halo exchange timed out (a neighbour never arrived); the wavefunction of this handle is invalid");
    return ION_OK;
}

int launch_observe(ion_sim *s, uint32_t what, double *dev_out)
{
    ion::ObserveParams p;
    std::memset(&p, 0, sizeof(p));
    p.psi = s->psi + (size_t)s->g_lo * s->Rp;
    p.rvec = s->rvec;
    p.h_diag = s->h_diag ? s->h_diag + (size_t)s->g_lo * s->R : nullptr;
    p.ghost_hi = s->g_hi;
    p.h_off = s->h_off;
    p.cl_z = s->cl_z;
    p.state_rows = s->state_rows;
    p.state_first = s->state_first;
    p.state_order = s->state_order;
    p.partial = s->partial;
    p.ip_out = s->ip_out;
    for (int q = 0; q < s->n_radii; ++q) p.radii[q] = s->radii[q];
    p.n_radii = (what & ION_OBS_NORM_WITHIN) ? s->n_radii : 0;
    p.n_states = (what & ION_OBS_INNER_PRODUCTS) ? s->n_states : 0;
    p.L = s->L_own;
    p.L_total = s->L_total;
    p.l_begin = s->l_own;
    p.R = s->R;
    p.M = s->M;
    p.T = s->T;
    p.what = what;
    p.ipm = s->ipm;
    p.line = s->line ? 1 : 0;
    prof_begin(s, KK_OBSERVE);
    ion::k_observe<<<dim3(s->L_own, s->batch), 256, 0, s->stream>>>(p);
    CUDA_TRY(cudaGetLastError());
    s->launch_count++;
    int rc = launch_observe_finish(s, what, dev_out);
    prof_end(s);
    return rc;
}

// record assembly from the per-channel partial sums and the inner products (left by k_observe or by a kernel with a fused observation)
int launch_observe_finish(ion_sim *s, uint32_t what, double *dev_out)
{
    const int n_states = (what & ION_OBS_INNER_PRODUCTS) ? s->n_states : 0, n_radii = (what & ION_OBS_NORM_WITHIN) ? s->n_radii : 0;
    const long long rec = ion_sim_observation_size(s, what);
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(s->batch);
    cfg.blockDim = dim3(256);
    cfg.stream = s->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = s->use_pdl ? 1 : 0;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, ion::k_observe_finish, (const double *)s->partial, (const double *)s->ip_out, dev_out, s->batch, s->L_own, n_states, n_radii,
                                (unsigned)what, s->ipm, rec));
    s->launch_count++;
    return ION_OK;
}

int ensure_observe_buffers(ion_sim *s, size_t n_records, uint32_t what)
{
    if (!s->partial) {
        if (int rc = dev_alloc(&s->partial, (size_t)s->batch * s->L_own * (4 + ION_MAX_RADII))) return rc;
    }
    if (!s->ip_out) {
        const size_t n = (size_t)s->batch * std::max(s->n_states, 1) * 2;
        if (int rc = dev_alloc(&s->ip_out, n)) return rc;
        CUDA_TRY(cudaMemset(s->ip_out, 0, n * sizeof(double)));  // states living in another l-block shard stay 0 here
    }
    size_t need = n_records * s->batch * (size_t)ion_sim_observation_size(s, what);
    if (need > s->obs_cap) {
        if (int rc = dev_alloc(&s->obs_out, need)) return rc;
        s->obs_cap = need;
    }
    return ION_OK;
}

// steps per captured graph: a chunk boundary costs a scalar refill (D2D copy) and a graph launch without overlap
const int64_t GRAPH_CHUNK = [] {
    if (const char *env = std::getenv("ION_GRAPH_CHUNK")) {
        const long v = std::atol(env);
        if (v >= 4 && v <= 4096) return (int64_t)v;
    }
    return (int64_t)64;
}();

// the main stream waits for the record assemblies still running on the side branch (they use partial / ip_out, which
// k_observe is about to overwrite; and a captured chunk must join everything it forked)
int join_side(ion_sim *s)
{
    for (int k = 0; k < 2; ++k)
        if (s->side_pending[k]) {
            CUDA_TRY(cudaStreamWaitEvent(s->stream, s->ev_done[k], 0));
            s->side_pending[k] = false;
        }
    return ION_OK;
}

// Can the observation of a step ride on the fused schedule?  Point-wise observables only (<z> and <H0> couple neighbouring
// channels / rows across CTA edges); velocity gauge with the inter-solve kernel, or the folded length-gauge step; one CTA per channel.
bool obs_fusable(const ion_sim *s, uint32_t what)
{
    const char *e = std::getenv("ION_NO_FUSED_OBS");  // A/B switch: every observed step on the single-sweep schedule + k_observe
    if ((e && e[0] == '1') || !what || s->S != 1 || s->M != 4 || s->L_own != s->L_total) return false;
    if (what & ~(ION_OBS_NORM | ION_OBS_INNER_PRODUCTS | ION_OBS_NORM_BY_L | ION_OBS_R | ION_OBS_NORM_WITHIN)) return false;
    if (s->program == ION_SH_VEL_SO) return s->slab_state == 1 && s->slab_partial && s->slab_ip;
    if (s->program == ION_SH_LEN_SO) return s->len_fold_state == 1;
    return false;
}

// enqueue steps [n0, n0+len) reading scalars from `scal` (row n - n0 of it) and writing observations to `obs`
int enqueue_steps(ion_sim *s, int64_t len, const double *scal, const uint8_t *pattern, bool pre_done_first, bool fuse_last,
                  uint32_t what, double *obs, size_t rec, bool *pre_done_out)
{
    const bool can_fuse = program_fuses(s);
    const bool fuse_obs = can_fuse && obs_fusable(s, what);
    const bool vel = s->program == ION_SH_VEL_SO;
    bool pre_done = pre_done_first;
    int64_t k_obs = 0;
    ObsReq pending;  // length gauge: the observation of the previous step, to be made by this step's kernel
    for (int64_t n = 0; n < len; ++n) {
        const bool ob = pattern && pattern[n];
        const bool last = n + 1 == len;
        // an observed step keeps the fused schedule when the reductions can ride along; the last step of a chunk hands a
        // finished state to whatever follows (the next chunk's graph, the caller), so its observation is a separate pass
        const bool fuse_next = can_fuse && (!ob || (fuse_obs && !last)) && (last ? fuse_last : true);
        const double *sa = scal + (size_t)n * s->batch;
        ObsReq cur;
        if (ob && fuse_next) {
            cur.what = what;
            cur.dst = obs + (size_t)k_obs * rec;
        }
        if (int rc = enqueue_step(s, sa, sa + s->batch, pre_done ? fuse_level(s) : 0, fuse_next, pending, vel ? cur : ObsReq())) return rc;
        pending = vel ? ObsReq() : cur;
        pre_done = fuse_next;
        if (ob) {
            if (!fuse_next) {
                if (int rc = join_side(s)) return rc;
                if (int rc = launch_observe(s, what, obs + (size_t)k_obs * rec)) return rc;
            }
            ++k_obs;
        }
    }
    if (pre_done_out) *pre_done_out = pre_done;
    if (int rc = join_side(s)) return rc;
    return restore_home(s);  // a captured chunk must end where it started
}

int run_impl(ion_sim *s, int64_t n_steps, const double *taus, const double *fields, const uint8_t *observe_mask, uint32_t what,
             double *out)
{
    if (n_steps < 0) return fail(ION_EINVAL, "n_steps < 0");
    if (n_steps == 0) return ION_OK;
    if (!taus || !fields) return fail(ION_EINVAL, "taus/fields must not be NULL");
    CUDA_TRY(cudaSetDevice(s->device));
    if (int rc = check_ready(s)) return rc;
    if (s->L_own != s->L_total && !s->peers_attached)
        return fail(ION_ESTATE, "an l-block shard is advanced with ion_sim_step_phase (caller-side halo exchange between phases) or, after "
                                "ion_sim_attach_peer, with ion_sim_step/run (peer-memory exchange inside the engine)");
    int64_t n_obs = 0;
    if (observe_mask)
        for (int64_t n = 0; n < n_steps; ++n) n_obs += observe_mask[n] ? 1 : 0;
    if (n_obs && !out) return fail(ION_EINVAL, "out must not be NULL when observations are requested");
    if (n_obs)
        if (int rc = ensure_observe_buffers(s, (size_t)n_obs, what)) return rc;
    if (int rc = upload_scalars(s, n_steps, taus, fields)) return rc;
    const size_t rec = (size_t)ion_sim_observation_size(s, what) * s->batch;
    const bool can_fuse = program_fuses(s);

    // tau must be constant (to 1e-9 relative: linspace jitter) for the cached LU factors; a changing
    // time step re-factors between steps and is run with plain launches.
    bool uniform_tau = true;
    for (int64_t n = 1; n < n_steps; ++n)
        if (std::fabs(taus[n] - taus[0]) > 1e-9 * std::fabs(taus[0])) uniform_tau = false;

    if (int rc = slab_prepare(s)) return rc;
    if (int rc = len_fold_prepare(s)) return rc;
    if (s->S > 1 || s->program == ION_SH_LEN_ADI)
        if (int rc = ensure_second_buffer(s)) return rc;  // segmented kernels run out of place
    if (n_obs && s->len_fold_state == 1 && s->S == 1 && s->L_own == s->L_total) {  // length gauge: the record assembly runs on a side branch
        if (!s->partial2)
            if (int rc = dev_alloc(&s->partial2, (size_t)s->batch * s->L_own * (4 + ION_MAX_RADII))) return rc;
        if (!s->ip_out2) {
            const size_t n = (size_t)s->batch * std::max(s->n_states, 1) * 2;
            if (int rc = dev_alloc(&s->ip_out2, n)) return rc;
            CUDA_TRY(cudaMemsetAsync(s->ip_out2, 0, n * sizeof(double), s->stream));
        }
        if (!s->side) {
            CUDA_TRY(cudaStreamCreateWithFlags(&s->side, cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
            for (auto &e : s->ev_done) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
    }
    if (n_obs && s->slab_state == 1) {  // opt-in: the observed state is stored and reduced on the side branch (slab.cuh: STORE)
        const char *env = std::getenv("ION_SLAB_OBS_STORE");
        const size_t n = (size_t)s->batch * s->L * s->Rp;
        // measured on C3: 47.1 us per observed step against 42.7 us with the in-kernel reductions (k_observe does not fit into the SM time the
        // pair kernel's second wave leaves idle), so this is opt-in (ION_SLAB_OBS_STORE=1; it carries every observable k_observe knows)
        if (env && env[0] == '1' && n * sizeof(cplx) <= ((size_t)256 << 20)) {
            for (auto &q : s->obs_psi)
                if (!q) {
                    if (int rc = dev_alloc(&q, n)) return rc;
                    CUDA_TRY(cudaMemsetAsync(q, 0, n * sizeof(cplx), s->stream));  // the padding rows are never written
                }
        }
    }
    if (n_obs && s->slab_state == 1) {  // per-slab partial sums of the fused observation (slab.cuh)
        s->slab_partial_half = (size_t)s->batch * s->slab_slabs * s->L * (4 + ION_MAX_RADII);
        s->slab_ip_half = (size_t)s->batch * s->slab_slabs * std::max(s->n_states, 1) * 2;
        if (!s->slab_partial)
            if (int rc = dev_alloc(&s->slab_partial, 2 * s->slab_partial_half)) return rc;
        if (!s->slab_ip)
            if (int rc = dev_alloc(&s->slab_ip, 2 * s->slab_ip_half)) return rc;
        if (!s->side) {
            CUDA_TRY(cudaStreamCreateWithFlags(&s->side, cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
            for (auto &e : s->ev_done) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
        if (!s->obs_counter) {
            if (int rc = dev_alloc(&s->obs_counter, (size_t)s->batch)) return rc;
            CUDA_TRY(cudaMemsetAsync(s->obs_counter, 0, (size_t)s->batch * sizeof(unsigned), s->stream));
        }
    }
    if (!(s->use_graphs && uniform_tau && !s->profiling && s->stream != 0 && n_steps >= 4)) {
        bool pre_done = false;
        int64_t k_obs = 0;
        for (int64_t n = 0; n < n_steps; ++n) {
            if (int rc = ensure_factor(s, taus[n])) return rc;
            const bool obs = observe_mask && observe_mask[n];
            const bool fuse_next = can_fuse && (n + 1 < n_steps) && !obs;
            const double *sa = s->scal + (size_t)(n + 1) * s->batch;
            if (int rc = enqueue_step(s, sa, sa + s->batch, pre_done ? fuse_level(s) : 0, fuse_next)) return rc;
            pre_done = fuse_next;
            if (obs) {
                if (int rc = launch_observe(s, what, s->obs_out + (size_t)k_obs * rec)) return rc;
                ++k_obs;
            }
        }
        if (int rc = restore_home(s)) return rc;
    } else {
        if (int rc = ensure_factor(s, taus[0])) return rc;
        if (!s->scal_chunk)
            if (int rc = dev_alloc(&s->scal_chunk, (size_t)(GRAPH_CHUNK + 2) * s->batch)) return rc;
        if (n_obs && s->obs_chunk_cap < (size_t)GRAPH_CHUNK * rec) {
            s->invalidate_graphs();
            if (int rc = dev_alloc(&s->obs_chunk, (size_t)GRAPH_CHUNK * rec)) return rc;
            s->obs_chunk_cap = (size_t)GRAPH_CHUNK * rec;
        }
        bool pre_done = false;
        int64_t k_obs = 0;
        for (int64_t n0 = 0; n0 < n_steps; n0 += GRAPH_CHUNK) {
            const int64_t len = std::min<int64_t>(GRAPH_CHUNK, n_steps - n0);
            const uint8_t *pattern = observe_mask ? observe_mask + n0 : nullptr;
            int64_t obs_here = 0;
            std::string key(1, (char)(pre_done ? 1 : 0));
            const bool fuse_last = can_fuse && (n0 + len < n_steps) && !(pattern && pattern[len - 1]);
            key.push_back((char)(fuse_last ? 1 : 0));
            key.append(reinterpret_cast<const char *>(&len), sizeof(len));
            key.append(reinterpret_cast<const char *>(&what), sizeof(what));
            for (int64_t n = 0; n < len; ++n) {
                const char ob = (pattern && pattern[n]) ? 1 : 0;
                obs_here += ob;
                key.push_back(ob);
            }
            auto it = s->graphs.find(key);
            bool pre_done_after = pre_done;
            if (it == s->graphs.end()) {
                ion_sim::GraphEntry entry;
                cudaGraph_t graph = nullptr;
                const int64_t before = s->launch_count;
                CUDA_TRY(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
                s->capturing = true;
                int rc = enqueue_steps(s, len, s->scal_chunk + s->batch, pattern, pre_done, fuse_last, what, s->obs_chunk, rec, &pre_done_after);
                s->capturing = false;
                cudaError_t e = cudaStreamEndCapture(s->stream, &graph);
                entry.launches = s->launch_count - before;
                s->launch_count = before;
                if (rc) {
                    if (graph) cudaGraphDestroy(graph);
                    return rc;
                }
                if (e != cudaSuccess) return fail(ION_ECUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
                e = cudaGraphInstantiate(&entry.exec, graph, 0);
                cudaGraphDestroy(graph);
                if (e != cudaSuccess) return fail(ION_ECUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
                it = s->graphs.emplace(key, entry).first;
            } else {
                // same bookkeeping as the capture run: only the last step of the chunk decides
                pre_done_after = can_fuse && fuse_last;
            }
            CUDA_TRY(cudaMemcpyAsync(s->scal_chunk, s->scal + (size_t)n0 * s->batch, (size_t)(len + 2) * s->batch * sizeof(double),
                                     cudaMemcpyDeviceToDevice, s->stream));
            CUDA_TRY(cudaGraphLaunch(it->second.exec, s->stream));
            s->launch_count += it->second.launches;
            if (obs_here) {
                CUDA_TRY(cudaMemcpyAsync(s->obs_out + (size_t)k_obs * rec, s->obs_chunk, (size_t)obs_here * rec * sizeof(double),
                                         cudaMemcpyDeviceToDevice, s->stream));
                k_obs += obs_here;
            }
            pre_done = pre_done_after;
        }
    }
    if (n_obs) {
        CUDA_TRY(cudaMemcpyAsync(out, s->obs_out, (size_t)n_obs * rec * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        if (int rc = check_abort(s)) return rc;
    }
    return ION_OK;
}

}  // namespace

// =============================================================================================
// C-ABI
// =============================================================================================
extern "C" {

int ion_abi_version(void) { return ION_ABI_VERSION; }
const char *ion_last_error(void) { return g_last_error.c_str(); }

int ion_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int ion_tdma_c128(const void *sub, const void *diag, const void *sup, const void *rhs, void *x, int64_t n, int64_t batch,
                  int device)
{
    if (n < 1 || batch < 0) return fail(ION_EINVAL, "ion_tdma_c128: need n >= 1, batch >= 0");
    if (batch == 0) return ION_OK;
    if (!diag || !rhs || !x || (n > 1 && (!sub || !sup))) return fail(ION_EINVAL, "ion_tdma_c128: NULL array");
    if (ion_device_count() <= device) return fail(ION_ENODEVICE, "ion_tdma_c128: no CUDA device " + std::to_string(device));
    CUDA_TRY(cudaSetDevice(device));
    cplx *d_sub = nullptr, *d_diag = nullptr, *d_sup = nullptr, *d_rhs = nullptr, *d_x = nullptr, *d_scratch = nullptr;
    const size_t nb = (size_t)batch * n * sizeof(cplx), nb1 = (size_t)batch * (n - 1) * sizeof(cplx);
    auto cleanup = [&]() {
        cudaFree(d_sub), cudaFree(d_diag), cudaFree(d_sup), cudaFree(d_rhs), cudaFree(d_x), cudaFree(d_scratch);
    };
    int rc = ION_OK;
    do {
        if ((rc = dev_alloc(&d_diag, (size_t)batch * n)) || (rc = dev_alloc(&d_rhs, (size_t)batch * n)) ||
            (rc = dev_alloc(&d_x, (size_t)batch * n)) || (rc = dev_alloc(&d_sub, (size_t)batch * std::max<int64_t>(n - 1, 1))) ||
            (rc = dev_alloc(&d_sup, (size_t)batch * std::max<int64_t>(n - 1, 1))))
            break;
        cudaError_t e = cudaSuccess;
        if (n > 1) {
            e = cudaMemcpy(d_sub, sub, nb1, cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = cudaMemcpy(d_sup, sup, nb1, cudaMemcpyHostToDevice);
        }
        if (e == cudaSuccess) e = cudaMemcpy(d_diag, diag, nb, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(d_rhs, rhs, nb, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            rc = fail(ION_ECUDA, std::string("ion_tdma_c128 H2D: ") + cudaGetErrorString(e));
            break;
        }
        size_t smem = 2 * (size_t)n * sizeof(cplx);
        if (smem > 160 * 1024) {
            if ((rc = dev_alloc(&d_scratch, (size_t)batch * 2 * n))) break;
            smem = 0;
        } else if (smem > 48 * 1024) {
            e = cudaFuncSetAttribute(ion::k_tdma_general, cudaFuncAttributeMaxDynamicSharedMemorySize, 164 * 1024);
        }
        if (e == cudaSuccess) {
            ion::k_tdma_general<<<(unsigned)batch, 128, smem>>>(d_sub, d_diag, d_sup, d_rhs, d_x, n, d_scratch);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaMemcpy(x, d_x, nb, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail(ION_ECUDA, std::string("ion_tdma_c128: ") + cudaGetErrorString(e));
    } while (0);
    cleanup();
    return rc;
}

int ion_sim_create_sharded(int program, int64_t L_total, int64_t l_begin, int64_t L, int64_t R, int64_t batch, int device,
                           ion_sim_t **out)
{
    if (!out) return fail(ION_EINVAL, "out is NULL");
    *out = nullptr;
    if (L < 1 || R < 2 || batch < 1 || l_begin < 0 || l_begin + L > L_total)
        return fail(ION_EINVAL, "need L >= 1, R >= 2, batch >= 1, 0 <= l_begin, l_begin + L <= L_total");
    const bool line = (program == ION_LINE_LEN_CN || program == ION_LINE_LEN_SO || program == ION_LINE_VEL_SO);
    if (program < 0 || program > ION_SH_LEN_ADI) return fail(ION_EINVAL, "unknown program");
    if (line && L_total != 1) return fail(ION_EINVAL, "LineMesh programs need L = 1");
    if (program == ION_SH_LEN_ADI && L_total > ion::ADI_CL * ion::ADI_MAX_THREADS)
        return fail(ION_ENOTSUP, "ION_SH_LEN_ADI: at most 4096 channels (one CTA holds all channels of a radial position)");
    const bool sharded = (L != L_total);
    int cut_par = 0;
    if (sharded) {
        if (batch != 1) return fail(ION_ENOTSUP, "l-block shards hold one simulation (batch = 1)");
        if (program != ION_SH_LEN_SO && program != ION_SH_VEL_SO) return fail(ION_ENOTSUP, "l-block sharding: split-operator SphericalHarmonic programs only");
        if ((L_total % 2) != 0) return fail(ION_ENOTSUP, "l-block sharding needs an even l_bound");
        const int64_t cut_lo = l_begin > 0 ? l_begin : -1, cut_hi = l_begin + L < L_total ? l_begin + L : -1;
        const int par_lo = cut_lo >= 0 ? (int)(cut_lo & 1) : -1, par_hi = cut_hi >= 0 ? (int)(cut_hi & 1) : -1;
        if (par_lo >= 0 && par_hi >= 0 && par_lo != par_hi) return fail(ION_EINVAL, "both cuts of an l-block shard must have the same parity");
        cut_par = par_lo >= 0 ? par_lo : par_hi;
        if (cut_par == 1 && program != ION_SH_LEN_SO)
            return fail(ION_EINVAL, "l-block shards of this program must begin and end on even channels (so that only odd sweeps cross a cut); "
                                    "odd cuts are for the length-gauge split-operator program");
    }
    if (ion_device_count() <= device || device < 0)
        return fail(ION_ENODEVICE, "no CUDA device " + std::to_string(device) + " (the engine has no CPU path)");
    int M = 4;
    if (const char *env = std::getenv("ION_M"))
    {
        if (env[0] == '8' && R <= 2048 && program != ION_SH_LEN_ADI) M = 8;              // experiment: 256 threads x 8 rows
    }
    // Long LineMesh channels of the Crank-Nicolson program (r-segments): EIGHT rows per thread.  The scans are the expensive
    // part -- the program rebuilds its pivots every step with a Kogge-Stone scan of 2x2 complex Moebius matrices -- and
    // their cost per thread does not depend on the rows a thread holds; the 256-row halos shrink from 64 to 32 threads, and at 128
    // registers two 256-thread CTAs share an SM.  configs[1] (1024 x 2^16): 1889 -> 1108 us per step (17.4 -> 29.6 % of the roofline).
    if (program == ION_LINE_LEN_CN && R > 4096) {
        const char *env = std::getenv("ION_LINE_M");
        if (!(env && env[0] == '4')) M = 8;
    }
    int64_t T = (R + M - 1) / M;
    T = (T + 31) / 32 * 32;
    int S = 1, T_seg = (int)T, H = 0;
    if (T > (M == 8 ? 256 : 1024)) {  // r-segments with halos (kernels.cuh); the halo width is fixed when the LU factors are built
        // + 2 x 32 halo threads per unit of reach (2 x 64 with r-pair bricks): 448 or 512 threads per CTA.  The length-gauge
        // split-operator step (32-thread halos) takes 192-thread segments instead: 256-thread CTAs, two per SM at 128 registers,
        // whose load / solve / store phases overlap -- 33 % recomputed rows instead of 17 %, and still 7 % faster on the
        // HBM-resident 16384 x 4096 mesh (1274 -> 1180 us per step; 160: 1275, 256: 1548, 320: 1346)
        // With half-warp halos (ensure_factor: the multipliers decay below 1e-18 over 64 rows) the same 256-thread CTA holds 224
        // interior threads: 14 % recomputed rows.
        T_seg = (program == ION_SH_LEN_SO) ? 224 : (M == 8 ? 192 : 384);
        if (const char *env = std::getenv("ION_TSEG")) {  // A/B switch
            const int v = std::atoi(env);
            if (v >= 64 && v <= 384 && v % 32 == 0) T_seg = v;
        }
        H = M == 8 ? 32 : 64;
        S = (int)((T + T_seg - 1) / T_seg);
        T = (int64_t)S * T_seg;
    }
    CUDA_TRY(cudaSetDevice(device));
    ion_sim *s = new ion_sim();
    s->program = program;
    s->g_lo = (sharded && l_begin > 0) ? 1 : 0;
    s->g_hi = (sharded && l_begin + L < L_total) ? 1 : 0;
    s->L_own = (int)L;
    s->l_own = (int)l_begin;
    s->cut_parity = cut_par;
    s->L = (int)L + s->g_lo + s->g_hi;
    s->l_begin = (int)l_begin - s->g_lo;
    L = s->L;
    s->R = (int)R;
    s->batch = (int)batch;
    s->device = device;
    s->L_total = (int)L_total;
    s->M = M;
    s->T = (int)T;
    s->S = S;
    s->T_seg = T_seg;
    s->H = H;
    s->Tc = T_seg + 2 * H;
    s->Rp = (int)(M * T);
    s->tmax = s->Tc <= 256 ? 256 : (s->Tc <= 512 ? 512 : 1024);
    s->line = line;
    if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) == cudaSuccess) s->own_stream = true;
    else s->stream = 0;
    if (const char *env = std::getenv("ION_NO_GRAPHS")) s->use_graphs = !(env[0] == '1');
    if (const char *env = std::getenv("ION_NO_PDL")) s->use_pdl = !(env[0] == '1');
    if (const char *env = std::getenv("ION_NO_SLAB")) s->use_slab = !(env[0] == '1');
    if (const char *env = std::getenv("ION_NO_LEN_FOLD")) s->use_len_fold = !(env[0] == '1');
    if (const char *env = std::getenv("ION_NO_ENS")) s->use_ens = !(env[0] == '1');
    int rc = prepare_kernels(s);
    if (rc == ION_OK) rc = dev_alloc(&s->psi, (size_t)batch * L * s->Rp);
    if (rc == ION_OK) {
        cudaError_t e = cudaMemset(s->psi, 0, (size_t)batch * L * s->Rp * sizeof(cplx));
        if (e != cudaSuccess) rc = fail(ION_ECUDA, cudaGetErrorString(e));
    }
    if (rc != ION_OK) {
        delete s;
        return rc;
    }
    *out = s;
    return ION_OK;
}

int ion_sim_create(int program, int64_t L, int64_t R, int64_t batch, int device, ion_sim_t **out)
{
    return ion_sim_create_sharded(program, L, 0, L, R, batch, device, out);
}

int ion_sim_destroy(ion_sim_t *s)
{
    if (!s) return ION_OK;
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    delete s;
    return ION_OK;
}

int ion_sim_set_stream(ion_sim_t *s, void *cuda_stream)
{
    if (!s) return fail(ION_EINVAL, "sim is NULL");
    s->invalidate_graphs();
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
    s->own_stream = false;
    s->stream = (cudaStream_t)cuda_stream;
    return ION_OK;
}

int ion_sim_set_hamiltonian(ion_sim_t *s, const void *h_diag, const double *h_off)
{
    if (!s || !h_diag || !h_off) return fail(ION_EINVAL, "NULL argument");
    s->invalidate_graphs();
    CUDA_TRY(cudaSetDevice(s->device));
    if (int rc = dev_alloc(&s->h_diag, (size_t)s->L * s->R)) return rc;
    if (int rc = dev_alloc(&s->h_off, (size_t)s->R - 1)) return rc;
    CUDA_TRY(cudaMemcpyAsync(s->h_diag, h_diag, (size_t)s->L * s->R * sizeof(cplx), cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaMemcpyAsync(s->h_off, h_off, ((size_t)s->R - 1) * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    s->h_off_host.assign(h_off, h_off + (s->R - 1));
    s->have_h = true;
    s->factored = false;
    return ION_OK;
}

int ion_sim_set_len_coupling(ion_sim_t *s, const double *c_l, const double *x_j)
{
    if (!s || !x_j || (s->L_total > 1 && !c_l)) return fail(ION_EINVAL, "NULL argument");
    s->invalidate_graphs();
    if (s->program != ION_SH_LEN_SO && s->program != ION_SH_LEN_ADI) return fail(ION_EINVAL, "length-gauge coupling does not belong to this program");
    CUDA_TRY(cudaSetDevice(s->device));
    if (int rc = upload_plain(s, c_l, (size_t)s->L_total - 1, &s->cl)) return rc;
    if (int rc = upload_plain(s, c_l, (size_t)s->L_total - 1, &s->cl_z)) return rc;
    if (int rc = upload_permuted(s, x_j, s->R, &s->vec, nullptr)) return rc;
    // x_j = -q r_j on the reference's uniform radial grid (meshes.py:1009-1011) is linear in j: the kernels then get the
    // rotation angles of a thread's consecutive rows by angle addition.  Verified here, to rounding, for whatever the caller passed.
    s->vec_dv = 0.0;
    if (s->R >= 3 && s->M == 4 && !std::getenv("ION_NO_LINEAR_ANGLES")) {
        const double dv = (x_j[s->R - 1] - x_j[0]) / (double)(s->R - 1);
        double dev = 0.0, mx = 0.0;
        for (int64_t j = 0; j < s->R; ++j) {
            dev = std::max(dev, std::fabs(x_j[j] - (x_j[0] + (double)j * dv)));
            mx = std::max(mx, std::fabs(x_j[j]));
        }
        if (dv != 0.0 && dev <= 8.0 * 2.220446049250313e-16 * mx) s->vec_dv = dv;
    }
    s->have_coupling = true;
    return ION_OK;
}

int ion_sim_set_vel_coupling(ion_sim_t *s, const double *c_l, const double *f1_l, const double *y_j, const double *z_j)
{
    if (!s || !y_j || !z_j || (s->L_total > 1 && (!c_l || !f1_l))) return fail(ION_EINVAL, "NULL argument");
    s->invalidate_graphs();
    if (s->program != ION_SH_VEL_SO) return fail(ION_EINVAL, "velocity-gauge coupling does not belong to this program");
    CUDA_TRY(cudaSetDevice(s->device));
    if (int rc = upload_plain(s, f1_l, (size_t)s->L_total - 1, &s->cl)) return rc;
    if (int rc = upload_plain(s, c_l, (size_t)s->L_total - 1, &s->cl2)) return rc;
    if (int rc = upload_plain(s, c_l, (size_t)s->L_total - 1, &s->cl_z)) return rc;
    if (int rc = upload_permuted(s, y_j, s->R, &s->vec, nullptr)) return rc;
    if (int rc = upload_permuted(s, z_j, s->R - 1, &s->zvec, &s->zprev)) return rc;
    s->have_coupling = true;
    return ION_OK;
}

int ion_sim_set_line_coupling(ion_sim_t *s, const double *w_z, double v_pref)
{
    if (!s) return fail(ION_EINVAL, "sim is NULL");
    s->invalidate_graphs();
    if (!s->line) return fail(ION_EINVAL, "line coupling does not belong to this program");
    CUDA_TRY(cudaSetDevice(s->device));
    if (s->program == ION_LINE_VEL_SO) {
        std::vector<double> z((size_t)s->R - 1, v_pref);
        if (int rc = upload_permuted(s, z.data(), s->R - 1, &s->zvec, &s->zprev)) return rc;
    } else {
        if (!w_z) return fail(ION_EINVAL, "w_z is NULL");
        if (int rc = upload_permuted(s, w_z, s->R, &s->vec, nullptr)) return rc;
    }
    s->have_coupling = true;
    return ION_OK;
}

int ion_sim_set_mask(ion_sim_t *s, const double *mask)
{
    if (!s) return fail(ION_EINVAL, "sim is NULL");
    s->invalidate_graphs();
    CUDA_TRY(cudaSetDevice(s->device));
    if (!mask) {
        if (s->mask) cudaFree(s->mask);
        s->mask = nullptr;
        return ION_OK;
    }
    return upload_permuted(s, mask, s->R, &s->mask, nullptr);
}

int ion_sim_set_observables(ion_sim_t *s, double ipm, const double *r_j, int64_t n_states, const int64_t *state_l,
                            const void *state_rows, int64_t n_radii, const double *radii)
{
    if (!s) return fail(ION_EINVAL, "sim is NULL");
    s->invalidate_graphs();
    if (n_radii > ION_MAX_RADII) return fail(ION_ENOTSUP, "at most 8 radii");
    if (n_states < 0 || n_radii < 0 || (n_states > 0 && (!state_l || !state_rows)) || (n_radii > 0 && !radii))
        return fail(ION_EINVAL, "inconsistent observables arguments");
    CUDA_TRY(cudaSetDevice(s->device));
    s->ipm = ipm;
    if (r_j) {
        if (int rc = upload_permuted(s, r_j, s->R, &s->rvec, nullptr)) return rc;
    }
    s->n_radii = (int)n_radii;
    for (int q = 0; q < n_radii; ++q) s->radii[q] = radii[q];
    // test states owned by this shard, grouped by channel (CSR)
    s->n_states = (int)n_states;
    if (s->ip_out) {
        cudaFree(s->ip_out);
        s->ip_out = nullptr;
    }
    if (s->ip_out2) {
        cudaFree(s->ip_out2);
        s->ip_out2 = nullptr;
    }
    if (s->slab_ip) {
        cudaFree(s->slab_ip);
        s->slab_ip = nullptr;
    }
    std::vector<int> first((size_t)s->L_own + 1, 0), order;
    for (int l = 0; l < s->L_own; ++l) {
        first[l] = (int)order.size();
        for (int64_t k = 0; k < n_states; ++k)
            if (state_l[k] == s->l_own + l) order.push_back((int)k);
    }
    first[s->L_own] = (int)order.size();
    for (int64_t k = 0; k < n_states; ++k)
        if (state_l[k] < 0 || state_l[k] >= s->L_total) return fail(ION_EINVAL, "state_l out of range");
    if (int rc = dev_alloc(&s->state_first, first.size())) return rc;
    CUDA_TRY(cudaMemcpy(s->state_first, first.data(), first.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (int rc = dev_alloc(&s->state_order, std::max<size_t>(order.size(), 1))) return rc;
    if (!order.empty())
        CUDA_TRY(cudaMemcpy(s->state_order, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (n_states > 0) {
        cplx *stage = nullptr;
        if (int rc = dev_alloc(&stage, (size_t)n_states * s->R)) return rc;
        if (int rc = dev_alloc(&s->state_rows, (size_t)n_states * s->Rp)) {
            cudaFree(stage);
            return rc;
        }
        cudaError_t e = cudaMemcpy(stage, state_rows, (size_t)n_states * s->R * sizeof(cplx), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) {
            dim3 grid((s->Rp + 127) / 128, (unsigned)std::min<int64_t>(n_states, 4096));
            ion::k_to_internal_c<<<grid, 128, 0, s->stream>>>(stage, s->state_rows, s->R, s->M, s->T, n_states, n_states);
            e = cudaStreamSynchronize(s->stream);
        }
        cudaFree(stage);
        if (e != cudaSuccess) return fail(ION_ECUDA, cudaGetErrorString(e));
    }
    return ION_OK;
}

int ion_sim_write_g(ion_sim_t *s, const void *g)
{
    if (!s || !g) return fail(ION_EINVAL, "NULL argument");
    CUDA_TRY(cudaSetDevice(s->device));
    const size_t n = (size_t)s->batch * s->L_own;  // sharded: batch == 1, ghosts are not part of g
    if (!s->io_stage)
        if (int rc = dev_alloc(&s->io_stage, n * s->R)) return rc;
    CUDA_TRY(cudaMemcpyAsync(s->io_stage, g, n * s->R * sizeof(cplx), cudaMemcpyHostToDevice, s->stream));
    dim3 grid((s->Rp + 127) / 128, (unsigned)std::min<size_t>(n, 8192));
    ion::k_to_internal_c<<<grid, 128, 0, s->stream>>>(s->io_stage, s->psi + (size_t)s->g_lo * s->Rp, s->R, s->M, s->T, (long long)n, (long long)n);
    s->launch_count++;
    CUDA_TRY(cudaGetLastError());
    return ION_OK;
}

int ion_sim_write_g_broadcast(ion_sim_t *s, const void *g)
{
    if (!s || !g) return fail(ION_EINVAL, "NULL argument");
    CUDA_TRY(cudaSetDevice(s->device));
    const size_t n = (size_t)s->batch * s->L_own, n1 = (size_t)s->L_own;
    if (!s->io_stage)
        if (int rc = dev_alloc(&s->io_stage, n * s->R)) return rc;
    CUDA_TRY(cudaMemcpyAsync(s->io_stage, g, n1 * s->R * sizeof(cplx), cudaMemcpyHostToDevice, s->stream));
    dim3 grid((s->Rp + 127) / 128, (unsigned)std::min<size_t>(n, 8192));
    ion::k_to_internal_c<<<grid, 128, 0, s->stream>>>(s->io_stage, s->psi + (size_t)s->g_lo * s->Rp, s->R, s->M, s->T, (long long)n, (long long)n1);
    s->launch_count++;
    CUDA_TRY(cudaGetLastError());
    return ION_OK;
}

int ion_sim_read_g(ion_sim_t *s, void *g)
{
    if (!s || !g) return fail(ION_EINVAL, "NULL argument");
    CUDA_TRY(cudaSetDevice(s->device));
    const size_t n = (size_t)s->batch * s->L_own;
    if (!s->io_stage)
        if (int rc = dev_alloc(&s->io_stage, n * s->R)) return rc;
    dim3 grid((s->Rp + 127) / 128, (unsigned)std::min<size_t>(n, 8192));
    ion::k_from_internal_c<<<grid, 128, 0, s->stream>>>(s->psi + (size_t)s->g_lo * s->Rp, s->io_stage, s->R, s->M, s->T, (long long)n);
    s->launch_count++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(g, s->io_stage, n * s->R * sizeof(cplx), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return check_abort(s);
}

int ion_sim_step(ion_sim_t *s, int64_t n_steps, const double *taus, const double *fields)
{
    if (!s) return fail(ION_EINVAL, "sim is NULL");
    return run_impl(s, n_steps, taus, fields, nullptr, 0, nullptr);
}

int64_t ion_sim_observation_size(ion_sim_t *s, uint32_t what)
{
    if (!s) return 0;
    int64_t n = 0;
    if (what & ION_OBS_NORM) n += 1;
    if (what & ION_OBS_INNER_PRODUCTS) n += 2 * (int64_t)s->n_states;
    if (what & ION_OBS_NORM_BY_L) n += s->L_own;
    if (what & ION_OBS_R) n += 1;
    if (what & ION_OBS_Z) n += 1;
    if (what & ION_OBS_H0) n += 1;
    if (what & ION_OBS_NORM_WITHIN) n += s->n_radii;
    return n;
}

int ion_sim_observe(ion_sim_t *s, uint32_t what, double *out)
{
    if (!s || !out) return fail(ION_EINVAL, "NULL argument");
    CUDA_TRY(cudaSetDevice(s->device));
    if (int rc = ensure_observe_buffers(s, 1, what)) return rc;
    if (int rc = launch_observe(s, what, s->obs_out)) return rc;
    const size_t rec = (size_t)ion_sim_observation_size(s, what) * s->batch;
    CUDA_TRY(cudaMemcpyAsync(out, s->obs_out, rec * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return check_abort(s);
}

int ion_sim_run(ion_sim_t *s, int64_t n_steps, const double *taus, const double *fields, const uint8_t *observe_mask,
                uint32_t what, double *out)
{
    if (!s) return fail(ION_EINVAL, "sim is NULL");
    int rc = run_impl(s, n_steps, taus, fields, observe_mask, what, out);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return check_abort(s);
}

int ion_sim_synchronize(ion_sim_t *s)
{
    if (!s) return fail(ION_EINVAL, "sim is NULL");
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return check_abort(s);
}

int ion_sim_halo_buffer(ion_sim_t *s, int which, void **device_ptr, int64_t *n_bytes)
{
    if (!s || !device_ptr || !n_bytes) return fail(ION_EINVAL, "NULL argument");
    if (which < 0 || which > 3) return fail(ION_EINVAL, "which must be 0..3");
    const size_t chan = (size_t)s->Rp;
    *n_bytes = (int64_t)(chan * sizeof(cplx));
    cplx *ptr = nullptr;
    switch (which) {
        case 0: ptr = s->psi + (size_t)s->g_lo * chan; break;                                   // send to lower: first owned
        case 1: ptr = s->psi + (size_t)(s->g_lo + s->L_own - 1) * chan; break;                  // send to upper: last owned
        case 2: ptr = s->g_lo ? s->psi : nullptr; break;                                        // recv from lower: ghost
        case 3: ptr = s->g_hi ? s->psi + (size_t)(s->g_lo + s->L_own) * chan : nullptr; break;  // recv from upper: ghost
    }
    *device_ptr = ptr;
    return ION_OK;
}


/* ---- peer-memory halo exchange between l-block shards (halo.cuh) ---- */
namespace {
struct PeerBlob {  // what a shard tells its neighbours (ION_PEER_BLOB_BYTES)
    cudaIpcMemHandle_t halo;  // 64 bytes: the halo block (flags + staging slots)
    int64_t channel_elems;    // Rp * batch
    int64_t has_lo, has_hi;   // neighbours this shard expects
    uint64_t local_halo;      // raw device pointer, valid inside the exporting process only
};
static_assert(sizeof(PeerBlob) == 96, "PeerBlob layout");

// flags, staging[2 sides][2 slots][Rp] of the exchange kernel, fstage[2 sides][2 slots][Rp] of the fused exchange (cplx = 2 words)
size_t halo_block_words(const ion_sim *s) { return (size_t)ion::HF_COUNT + 2 * (2 * 2 * (size_t)s->Rp * 2); }

int ensure_halo_flags(ion_sim *s)
{
    if (s->hflags) return ION_OK;
    if (int rc = dev_alloc(&s->hflags, halo_block_words(s))) return rc;
    CUDA_TRY(cudaMemset(s->hflags, 0, halo_block_words(s) * sizeof(unsigned long long)));
    if (int rc = dev_alloc(&s->hf_sent, 2 * (size_t)s->S)) return rc;
    CUDA_TRY(cudaMemset(s->hf_sent, 0, 2 * (size_t)s->S * sizeof(unsigned long long)));
    return ION_OK;
}
}  // namespace

int ion_sim_export_peer(ion_sim_t *s, void *blob, int64_t blob_bytes)
{
    if (!s || !blob) return fail(ION_EINVAL, "NULL argument");
    if (blob_bytes < (int64_t)sizeof(PeerBlob)) return fail(ION_EINVAL, "blob too small (ION_PEER_BLOB_BYTES)");
    if (s->L_own == s->L_total) return fail(ION_ESTATE, "not an l-block shard");
    CUDA_TRY(cudaSetDevice(s->device));
    if (int rc = ensure_halo_flags(s)) return rc;
    PeerBlob b;
    std::memset(&b, 0, sizeof(b));
    CUDA_TRY(cudaIpcGetMemHandle(&b.halo, s->hflags));
    b.channel_elems = (int64_t)s->Rp;
    b.has_lo = s->g_lo;
    b.has_hi = s->g_hi;
    b.local_halo = (uint64_t)(uintptr_t)s->hflags;
    std::memcpy(blob, &b, sizeof(b));
    return ION_OK;
}

int ion_sim_attach_peer(ion_sim_t *s, int side, const void *blob, int64_t blob_bytes, int same_process)
{
    if (!s || !blob) return fail(ION_EINVAL, "NULL argument");
    if (side < 0 || side > 1) return fail(ION_EINVAL, "side must be 0 (lower neighbour) or 1 (upper neighbour)");
    if (blob_bytes < (int64_t)sizeof(PeerBlob)) return fail(ION_EINVAL, "blob too small (ION_PEER_BLOB_BYTES)");
    if ((side == 0 && !s->g_lo) || (side == 1 && !s->g_hi)) return fail(ION_ESTATE, "this shard has no neighbour on that side");
    CUDA_TRY(cudaSetDevice(s->device));
    if (int rc = ensure_halo_flags(s)) return rc;
    PeerBlob b;
    std::memcpy(&b, blob, sizeof(b));
    if (b.channel_elems != (int64_t)s->Rp) return fail(ION_EINVAL, "the neighbour's radial layout differs from this shard's");
    // my lower neighbour receives my first channel on ITS upper side, my upper neighbour my last channel on its lower side
    if ((side == 0 && !b.has_hi) || (side == 1 && !b.has_lo)) return fail(ION_EINVAL, "the neighbour has no ghost channel facing this shard");
    unsigned long long *pblock = nullptr;
    if (same_process) {
        pblock = reinterpret_cast<unsigned long long *>((uintptr_t)b.local_halo);
        // several shards of ONE process on different GPUs (mesh API: SphericalHarmonicSpecification(devices=[...])): the
        // neighbour's halo block is dereferenced by this device's kernels, which needs peer access between the two devices
        cudaPointerAttributes pa;
        const cudaError_t pe = cudaPointerGetAttributes(&pa, pblock);
        if (pe != cudaSuccess || pa.type != cudaMemoryTypeDevice || pa.device == s->device) s->neighbour_on_same_device = true;
        if (pe == cudaSuccess && pa.type == cudaMemoryTypeDevice && pa.device != s->device) {
            int can = 0;
            CUDA_TRY(cudaDeviceCanAccessPeer(&can, s->device, pa.device));
            if (!can) return fail(ION_ENOTSUP, "devices " + std::to_string(s->device) + " and " + std::to_string(pa.device) + " cannot access each other's memory (no NVLink / PCIe peer path)");
            cudaError_t e = cudaDeviceEnablePeerAccess(pa.device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) return fail(ION_ECUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
        } else {
            cudaGetLastError();
        }
    } else {
        void *q = nullptr;
        CUDA_TRY(cudaIpcOpenMemHandle(&q, b.halo, cudaIpcMemLazyEnablePeerAccess));
        s->peer_ipc_base[side] = q;
        pblock = static_cast<unsigned long long *>(q);
    }
    const int facing = 1 - side;  // the neighbour's side that faces me
    s->peer_flags[side] = pblock;
    s->peer_stage[side] = reinterpret_cast<cplx *>(pblock + ion::HF_COUNT) + (size_t)facing * 2 * s->Rp;
    s->peer_fstage[side] = reinterpret_cast<cplx *>(pblock + ion::HF_COUNT) + (size_t)(4 + facing * 2) * s->Rp;
    s->peers_attached = (!s->g_lo || s->peer_flags[0]) && (!s->g_hi || s->peer_flags[1]);
    if (s->len_fold_state == -1) s->len_fold_state = 0;  // the folded length-gauge step of a shard needs the engine's own exchange: decide again
    s->invalidate_graphs();
    if (!s->side) {  // side branch for the exchange and the two boundary units of every odd-parity kernel (launch_odd_exchanged)
        CUDA_TRY(cudaStreamCreateWithFlags(&s->side, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
        for (auto &e : s->ev_done) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    return ION_OK;
}

int ion_sim_exchange_halos(ion_sim_t *s)
{
    if (!s) return fail(ION_EINVAL, "sim is NULL");
    if (!s->peers_attached) return fail(ION_ESTATE, "ion_sim_attach_peer has not been called for every neighbour");
    CUDA_TRY(cudaSetDevice(s->device));
    return launch_exchange(s);
}

int ion_sim_prepare(ion_sim_t *s, double tau)
{
    if (!s) return fail(ION_EINVAL, "sim is NULL");
    CUDA_TRY(cudaSetDevice(s->device));
    if (int rc = check_ready(s)) return rc;
    if (int rc = ensure_factor(s, tau)) return rc;
    if (s->S > 1 || s->program == ION_SH_LEN_ADI)
        if (int rc = ensure_second_buffer(s)) return rc;
    // linked shards must not allocate inside the step loop: decide the fused schedules (and their second buffer) here
    if (s->L_own == s->L_total || s->peers_attached) {
        if (int rc = slab_prepare(s)) return rc;
        if (int rc = len_fold_prepare(s)) return rc;
    }
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return ION_OK;
}

int ion_sim_reserve(ion_sim_t *s, int64_t n_steps, int64_t n_records, uint32_t what)
{
    if (!s) return fail(ION_EINVAL, "sim is NULL");
    if (n_steps < 0 || n_records < 0) return fail(ION_EINVAL, "negative count");
    CUDA_TRY(cudaSetDevice(s->device));
    // everything ion_sim_step / ion_sim_run would (re)allocate for a call of this size.  cudaFree waits for the device to go
    // idle, which never happens while another shard's halo kernel is waiting for THIS shard's kernels: linked shards must not
    // (re)allocate between hand-shakes, so their buffers are sized up front.
    const size_t need = (size_t)(n_steps + 2) * s->batch;
    if (need > s->scal_cap) {
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        if (int rc = dev_alloc(&s->scal, need)) return rc;
        s->scal_cap = need;
    }
    for (int k = 0; k < 2; ++k) {
        if (need > s->scal_host_cap[k]) {
            if (s->scal_ev[k]) CUDA_TRY(cudaEventSynchronize(s->scal_ev[k]));
            if (s->scal_host[k]) cudaFreeHost(s->scal_host[k]);
            s->scal_host[k] = nullptr;
            s->scal_host_cap[k] = 0;
            CUDA_TRY(cudaMallocHost((void **)&s->scal_host[k], need * sizeof(double)));
            s->scal_host_cap[k] = need;
        }
    }
    if (!s->scal_chunk)
        if (int rc = dev_alloc(&s->scal_chunk, (size_t)(GRAPH_CHUNK + 2) * s->batch)) return rc;
    if (n_records > 0) {
        if (int rc = ensure_observe_buffers(s, (size_t)n_records, what)) return rc;
        const size_t rec = (size_t)ion_sim_observation_size(s, what) * s->batch;
        if (s->obs_chunk_cap < (size_t)GRAPH_CHUNK * rec) {
            s->invalidate_graphs();
            if (int rc = dev_alloc(&s->obs_chunk, (size_t)GRAPH_CHUNK * rec)) return rc;
            s->obs_chunk_cap = (size_t)GRAPH_CHUNK * rec;
        }
    }
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return ION_OK;
}

int ion_sim_halo_status(ion_sim_t *s, int64_t *exchanges_done, int *aborted)
{
    if (!s || !exchanges_done || !aborted) return fail(ION_EINVAL, "NULL argument");
    *exchanges_done = 0;
    *aborted = 0;
    if (!s->hflags) return ION_OK;
    unsigned long long h[ion::HF_COUNT];
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaMemcpy(h, s->hflags, sizeof(h), cudaMemcpyDeviceToHost));
    *exchanges_done = (int64_t)h[ion::HF_SEQ];
    *aborted = h[ion::HF_ABORT] != 0ull;
    return ION_OK;
}

namespace {
struct Phase {
    int prog, parity, flags;
};
// one UNFUSED step as a list of pair-local kernels (same order as enqueue_step)
int phases_of(const ion_sim *s, Phase *out)
{
    using namespace ion;
    if (s->program == ION_SH_LEN_SO) {
        out[0] = {PROG_ROT, 0, 0};
        out[1] = {PROG_ROT_CN_ROT, 1, 0};
        out[2] = {PROG_ROT, 0, F_MASK};
        return 3;
    }
    if (s->program == ION_SH_VEL_SO) {
        out[0] = {PROG_ROT, 0, F_REAL_ROT};
        out[1] = {PROG_ROT, 1, F_REAL_ROT};
        out[2] = {PROG_H2, 0, 0};
        out[3] = {PROG_H2_CN_H2, 1, 0};
        out[4] = {PROG_H2, 0, F_H2_REVERSE};
        out[5] = {PROG_ROT, 1, F_REAL_ROT};
        out[6] = {PROG_ROT, 0, F_REAL_ROT | F_MASK};
        return 7;
    }
    return 0;
}
}  // namespace

int ion_sim_num_phases(ion_sim_t *s)
{
    if (!s) return 0;
    Phase ph[8];
    return phases_of(s, ph);
}

int ion_sim_phase_needs_halo(ion_sim_t *s, int phase)
{
    if (!s) return 0;
    Phase ph[8];
    const int n = phases_of(s, ph);
    if (phase < 0 || phase >= n) return 0;
    return ph[phase].parity != s->cut_parity ? 1 : 0;  // pairs (l, l+1) with l % 2 == parity straddle a cut c when c % 2 != parity
}

int ion_sim_step_phase(ion_sim_t *s, int phase, double tau, const double *field)
{
    if (!s || !field) return fail(ION_EINVAL, "NULL argument");
    CUDA_TRY(cudaSetDevice(s->device));
    if (int rc = check_ready(s)) return rc;
    Phase ph[8];
    const int n = phases_of(s, ph);
    if (n == 0) return fail(ION_ENOTSUP, "program has no phase decomposition");
    if (phase < 0 || phase >= n) return fail(ION_EINVAL, "phase out of range");
    if (int rc = ensure_factor(s, tau)) return rc;
    if (!s->scal_phase)
        if (int rc = dev_alloc(&s->scal_phase, (size_t)s->batch)) return rc;
    if (phase == 0) {
        std::vector<double> h((size_t)s->batch);
        for (int b = 0; b < s->batch; ++b) h[b] = tau * field[b];
        CUDA_TRY(cudaMemcpyAsync(s->scal_phase, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
    }
    if (s->S > 1 || s->program == ION_SH_LEN_ADI)
        if (int rc = ensure_second_buffer(s)) return rc;
    if (int rc = launch_unit(s, ph[phase].prog, ph[phase].parity, ph[phase].flags, s->scal_phase, nullptr)) return rc;
    return restore_home(s);  // the caller's halo buffers point into the home buffer
}

int ion_sim_device_psi(ion_sim_t *s, void **device_ptr, int64_t *n_bytes)
{
    if (!s || !device_ptr || !n_bytes) return fail(ION_EINVAL, "NULL argument");
    *device_ptr = s->psi;
    *n_bytes = (int64_t)((size_t)s->batch * s->L * s->Rp * sizeof(cplx));
    return ION_OK;
}

int ion_sinc_pulse_fields(int device, int kind, int64_t n_times, const double *times, double t_offset, int64_t n_pulses, const double *pulse_params,
                          double *out)
{
    if (!times || !pulse_params || !out) return fail(ION_EINVAL, "NULL argument");
    if (n_times < 2 || n_pulses < 1) return fail(ION_EINVAL, "need n_times >= 2 and n_pulses >= 1");
    if (kind != ION_FIELD_E && kind != ION_FIELD_A) return fail(ION_EINVAL, "kind must be ION_FIELD_E or ION_FIELD_A");
    if (ion_device_count() <= device || device < 0) return fail(ION_ENODEVICE, "no CUDA device " + std::to_string(device));
    CUDA_TRY(cudaSetDevice(device));
    double *d_t = nullptr, *d_e = nullptr, *d_out = nullptr;
    ion::SincPulseParams *d_p = nullptr;
    auto cleanup = [&]() { cudaFree(d_t), cudaFree(d_e), cudaFree(d_out), cudaFree(d_p); };
    const size_t n_e = (size_t)n_times * n_pulses, n_out = (size_t)(n_times - 1) * n_pulses;
    int rc = ION_OK;
    do {
        if ((rc = dev_alloc(&d_t, (size_t)n_times)) || (rc = dev_alloc(&d_e, n_e)) || (rc = dev_alloc(&d_p, (size_t)n_pulses)) ||
            (kind == ION_FIELD_A && (rc = dev_alloc(&d_out, n_out))))
            break;
        cudaError_t e = cudaMemcpy(d_t, times, (size_t)n_times * sizeof(double), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(d_p, pulse_params, (size_t)n_pulses * sizeof(ion::SincPulseParams), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) {
            // E: the samples the step n -> n + 1 uses are times[1..] + offset; A: samples at times[0..] (offset 0), then every prefix
            ion::k_sinc_field<<<dim3((unsigned)((n_pulses + 127) / 128), (unsigned)n_times), 128>>>(d_t, kind == ION_FIELD_E ? t_offset : 0.0, (int)n_times, d_p, (int)n_pulses, d_e);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess && kind == ION_FIELD_A) {
            ion::k_prefix_simps<<<(unsigned)((n_pulses + 127) / 128), 128>>>(d_e, d_t, (int)n_times, (int)n_pulses, -1.0, d_out);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess)
            e = cudaMemcpy(out, kind == ION_FIELD_A ? d_out : d_e + n_pulses, n_out * sizeof(double), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail(ION_ECUDA, std::string("ion_sinc_pulse_fields: ") + cudaGetErrorString(e));
    } while (0);
    cleanup();
    return rc;
}

int64_t ion_sim_launch_count(ion_sim_t *s) { return s ? s->launch_count : 0; }

// FP64 pipe peak, measured: independent DFMA chains at full occupancy (the second bound of the roofline: the hot path is
// complex128 arithmetic, no contraction, so the tensor cores do not apply and the FP64 FMA pipe is the compute ceiling)
namespace {
__global__ void __launch_bounds__(512) k_fp64_peak(double *out, int iters)
{
    double a = threadIdx.x * 1e-3, b = 1.0000001, c = 1e-9, d = a + 1, e = a + 2, f = a + 3, g = a + 4, h = a + 5, i2 = a + 6, j = a + 7;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        a = fma(a, b, c);
        d = fma(d, b, c);
        e = fma(e, b, c);
        f = fma(f, b, c);
        g = fma(g, b, c);
        h = fma(h, b, c);
        i2 = fma(i2, b, c);
        j = fma(j, b, c);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a + d + e + f + g + h + i2 + j;
}
}  // namespace

int ion_fp64_peak(int device, double *fma_per_second)
{
    if (!fma_per_second) return fail(ION_EINVAL, "NULL argument");
    if (ion_device_count() <= device || device < 0) return fail(ION_ENODEVICE, "no CUDA device " + std::to_string(device));
    CUDA_TRY(cudaSetDevice(device));
    int sms = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int grid = sms * 4, block = 512, iters = 8192;
    double *out = nullptr;
    CUDA_TRY(cudaMalloc((void **)&out, (size_t)grid * block * sizeof(double)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        k_fp64_peak<<<grid, block>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (e != cudaSuccess) return fail(ION_ECUDA, cudaGetErrorString(e));
    *fma_per_second = (double)grid * block * iters * 8.0 / (best * 1e-3);
    return ION_OK;
}
int ion_num_kernel_kinds(void) { return KK_COUNT; }
const char *ion_kernel_name(int kind) { return (kind >= 0 && kind < KK_COUNT) ? kKernelNames[kind] : ""; }

int ion_sim_profile(ion_sim_t *s, int64_t n_steps, const double *taus, const double *fields, double *ms, int64_t *launches)
{
    if (!s || !ms || !launches) return fail(ION_EINVAL, "NULL argument");
    for (int k = 0; k < KK_COUNT; ++k) ms[k] = 0.0, launches[k] = 0;
    s->profiling = true;
    int rc = run_impl(s, n_steps, taus, fields, nullptr, 0, nullptr);
    s->profiling = false;
    cudaError_t e = cudaStreamSynchronize(s->stream);
    if (rc == ION_OK && e == cudaSuccess) {
        for (size_t i = 0; i + 1 < s->ev.size(); i += 2) {
            float t = 0.f;
            cudaEventElapsedTime(&t, s->ev[i], s->ev[i + 1]);
            int k = s->ev_kind[i];
            if (k >= 0 && k < KK_COUNT) ms[k] += t, launches[k]++;
        }
    }
    for (auto ev : s->ev) cudaEventDestroy(ev);
    s->ev.clear();
    s->ev_kind.clear();
    if (rc) return rc;
    if (e != cudaSuccess) return fail(ION_ECUDA, cudaGetErrorString(e));
    return ION_OK;
}

}  // extern "C"
