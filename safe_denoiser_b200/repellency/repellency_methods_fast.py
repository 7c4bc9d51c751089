"""Drop-in for /root/reference/repellency/repellency_methods_fast.py (used by run_copro.py:52).

Same registry, constructor call, method names, return dicts and proj_ref cache format; the
[Q,N,D+1] torch broadcast (:249-250) is replaced by the CUDA projection.
"""
from ._base import make_registry
from ._kernel_family import FastRepellencyMethod, KernelFast, RandomNoise, Sparse, _unconstructable

__CONDITIONING_METHOD__, register_conditioning_method, get_repellency_method = make_registry()


class RepellencyMethod(FastRepellencyMethod):
    pass


@register_conditioning_method(name='euclidean')
class EuclideanRepellency(_unconstructable('euclidean')):
    pass


@register_conditioning_method(name='kernel')
class RBFKernelRepellencyLegacy(_unconstructable('kernel')):
    pass


@register_conditioning_method(name='kernel_fast')
class RBFKernelRepellency(KernelFast, RepellencyMethod):
    pass


@register_conditioning_method(name='random_noise')
class RandomNoiseRepellency(RandomNoise, RepellencyMethod):
    pass


@register_conditioning_method(name='sparse')
class SparseRepellency(Sparse, RepellencyMethod):
    pass


@register_conditioning_method(name='lsh')
class LSHRepellency(_unconstructable('lsh')):
    pass
