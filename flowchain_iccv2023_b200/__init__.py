"""flowchain_iccv2023_b200 -- B200-native analytical-density hot path of FlowChain (ICCV 2023).

Only what the path needs: csrc/ (sm_100a kernels + C-ABI, built into libflowchain_b200.so),
the host-side mirror of the reference flow / wrapper interface, weight packing and the
multi-GPU sharding helper.  See DESIGN.md and include/flowchain_b200.h.
"""
from .synth import FlowSpec, DEFAULT_SPEC  # noqa: F401


def __getattr__(name):
    # torch-dependent parts are imported lazily so that `import flowchain_iccv2023_b200.synth`
    # stays cheap for the CPU oracle tests.
    if name in ("fastpredNF_CIF_separate_cond", "CIF_step", "LinearMaskedCoupling", "BatchNorm", "FlowSequential",
                "DiagonalGaussianConditionalDensity"):
        from . import flow
        return getattr(flow, name)
    if name == "fastpredNF_TP":
        from .model import fastpredNF_TP
        return fastpredNF_TP
    raise AttributeError(name)
