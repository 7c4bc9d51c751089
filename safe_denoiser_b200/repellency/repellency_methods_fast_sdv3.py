"""Drop-in for /root/reference/repellency/repellency_methods_fast_sdv3.py (run_nudity_sdv3.py:32,
run_coco30k_sdv3.py:39): the fast module plus per-pixel channel L2-normalisation of the query
before the distance (:239).  The correction is still subtracted from the un-normalised x0 (Q7).
"""
from ._base import make_registry
from ._kernel_family import FastRepellencyMethod, KernelFast, RandomNoise, Sparse, _unconstructable

__CONDITIONING_METHOD__, register_conditioning_method, get_repellency_method = make_registry()


class RepellencyMethod(FastRepellencyMethod):
    normalize_query = True


@register_conditioning_method(name='euclidean')
class EuclideanRepellency(_unconstructable('euclidean')):
    pass


@register_conditioning_method(name='kernel')
class RBFKernelRepellencyLegacy(_unconstructable('kernel')):
    pass


@register_conditioning_method(name='kernel_fast')
class RBFKernelRepellency(KernelFast, RepellencyMethod):
    normalize_query = True


@register_conditioning_method(name='random_noise')
class RandomNoiseRepellency(RandomNoise, RepellencyMethod):
    pass


@register_conditioning_method(name='sparse')
class SparseRepellency(Sparse, RepellencyMethod):
    # fast_sdv3.py:332: distances and force are taken on the channel-normalised query, the update lands on x0
    normalize_query = True


@register_conditioning_method(name='lsh')
class LSHRepellency(_unconstructable('lsh')):
    pass
