"""Evolution-method strategy objects (ionization/mesh/evolution_methods.py).  They select the engine program; the
operator sequence itself is implemented by the CUDA kernels (ionization_b200/csrc/engine.cu: enqueue_step)."""


class EvolutionMethod:
    kind = None

    def __repr__(self):
        return f"{self.__class__.__name__}()"

    def info(self):
        return self.__class__.__name__

    def evolve(self, mesh, g, time_step):
        """EvolutionMethod.evolve(mesh, g, time_step) -> g (evolution_methods.py:19-24): one step WITHOUT the mask,
        on the device, for callers that drive the mesh by hand.  ``g`` must be the mesh's current wavefunction."""
        return mesh._evolve_operator_only(g, time_step)


class AlternatingDirectionImplicit(EvolutionMethod):
    """Crank-Nicolson / ADI (evolution_methods.py:46-77)"""

    kind = "adi"


class SplitInteractionOperator(EvolutionMethod):
    """split-operator (evolution_methods.py:80-123)"""

    kind = "so"
