"""Import helper for the UNMODIFIED reference in the build container (never used on the GPU box, never by the product).

The reference's HAT file imports three arithmetic-free symbols from the un-vendored, unpinned `basicsr` package
(requirements_hat.txt:11): the ARCH_REGISTRY decorator, to_2tuple and trunc_normal_ (models/hat_arch/hat_arch.py:6-7).
`install()` registers stand-ins in sys.modules so that the reference source itself is imported as is.
"""
import os
import sys
import types

REF = os.environ.get("SR_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "models"))


def install():
    import torch
    if "basicsr" not in sys.modules:
        basicsr = types.ModuleType("basicsr")
        utils = types.ModuleType("basicsr.utils")
        registry = types.ModuleType("basicsr.utils.registry")
        archs = types.ModuleType("basicsr.archs")
        arch_util = types.ModuleType("basicsr.archs.arch_util")

        class _Registry:
            def register(self, obj=None):
                return (lambda o: o) if obj is None else obj

        registry.ARCH_REGISTRY = _Registry()
        arch_util.to_2tuple = lambda v: tuple(v) if isinstance(v, (tuple, list)) else (v, v)
        arch_util.trunc_normal_ = torch.nn.init.trunc_normal_
        for name, mod in (("basicsr", basicsr), ("basicsr.utils", utils), ("basicsr.utils.registry", registry),
                          ("basicsr.archs", archs), ("basicsr.archs.arch_util", arch_util)):
            sys.modules[name] = mod
    if REF not in sys.path:
        sys.path.insert(0, REF)


def hat_module():
    """models/hat_arch/hat_arch.py — the file hybridmodels_hat.py actually imports."""
    install()
    import importlib.util
    path = os.path.join(REF, "models", "hat_arch", "hat_arch.py")
    spec = importlib.util.spec_from_file_location("ref_hat_arch", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def hybrid_module():
    install()
    import models.hybridmodels_hat as m
    return m
