"""Import-only cupy shim (pwc/correlation/correlation.py:5 imports cupy at module import)."""


def memoize(for_each_device=False):
    def deco(fn):
        return fn
    return deco


class _Cuda:
    @staticmethod
    def compile_with_cache(src):
        raise NotImplementedError("cupy is not available; kernels run under tests/golden/cuda_emu.py")


cuda = _Cuda()
