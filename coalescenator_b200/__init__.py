"""coalescenator_b200 — B200-native importance-sampling likelihood path of Coalescenator.

Only what the hot path needs lives here: csrc/ (CUDA kernels + the C ABI of include/coalgpu.h),
the host-side mirror of the reference's sampler interface (sampler.py), the locus-pair data model
and synthetic generator (problem.py) and the multi-GPU combine (distributed.py)."""
from .problem import Problem  # noqa: F401
from .sampler import ImportanceSampler, DeviceSampler, SamplerOutput, CoalGpuError, set_emission_mode  # noqa: F401
