"""Import the unmodified reference (compressionOrg/GRASP) from /root/reference.

Only usable where the reference is mounted (the build container); never on the GPU box.
`grasp.py` of the reference imports lm_eval (absent here) through evaluate_grasp.py, so a
three-module stub is installed first -- the hot path never touches it.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GRASP_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "modeling_grasp.py"))


def _stub_lm_eval():
    if "lm_eval" in sys.modules:
        return
    lm_eval = types.ModuleType("lm_eval")
    base = types.ModuleType("lm_eval.base")
    base.BaseLM = type("BaseLM", (), {})
    evaluator = types.ModuleType("lm_eval.evaluator")
    lm_eval.base, lm_eval.evaluator = base, evaluator
    sys.modules.update({"lm_eval": lm_eval, "lm_eval.base": base, "lm_eval.evaluator": evaluator})


class _RefModules:
    """Loads the reference's modules by file path under private names, so they never shadow the
    repo's own modeling_grasp / tools.utils_func (same file names by design)."""

    def __init__(self):
        if not available():
            raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
        import importlib.util

        def load_file(private_name, rel):
            spec = importlib.util.spec_from_file_location(private_name, os.path.join(REFERENCE_ROOT, rel))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[private_name] = mod
            spec.loader.exec_module(mod)
            return mod

        self.utils_func = load_file("_grasp_reference_utils_func", "tools/utils_func.py")
        # the reference's modeling_grasp does `from tools.utils_func import ...`: point that name at the
        # reference's helper while it is being imported, then restore whatever was there
        names = ("tools", "tools.utils_func")
        saved = {k: sys.modules.get(k) for k in names}
        pkg = types.ModuleType("tools")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "tools")]
        pkg.utils_func = self.utils_func
        sys.modules["tools"], sys.modules["tools.utils_func"] = pkg, self.utils_func
        try:
            self.modeling = load_file("_grasp_reference_modeling_grasp", "modeling_grasp.py")
        finally:
            for k in names:
                if saved[k] is not None:
                    sys.modules[k] = saved[k]
                else:
                    sys.modules.pop(k, None)


_CACHE = None


def load() -> _RefModules:
    global _CACHE
    if _CACHE is None:
        _CACHE = _RefModules()
    return _CACHE
