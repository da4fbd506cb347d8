"""Make the reference's own entry scripts (`Test.py`, `Demo.py`) run on the B200 backend UNCHANGED.

    import yoho_b200.dropin; yoho_b200.dropin.install()     # before `import tests.evaluator`

`install()` registers this package's modules under the names the reference imports —
`utils.network`, `utils.knn_search`, `tests.extractor`, `tests.matcher`, `tests.estimator` — so
`tests/evaluator.py:22-24` and `Demo.py:6-12` pick up the B200 implementations through the reference's own
`name2*` registries.  Everything else (`parses`, `utils.dataset`, `utils.RR_cal`, `tests.evaluator`) stays the
reference's.  `install(metrics=True)` additionally patches the numeric entry points of the reference's `utils.RR_cal`
(`rotation_error`, `translation_error`, `computeTransformationErr`, `evaluate_registration`) with `yoho_b200.rr_cal`, which
also removes that module's run-time need for `nibabel`.  See INTEGRATION.md.
"""
import importlib
import sys

_ALIASES = {
    "utils.network": "yoho_b200.network",
    "utils.knn_search": "yoho_b200.knn_search",
    "tests.extractor": "yoho_b200.extractor",
    "tests.matcher": "yoho_b200.matcher",
    "tests.estimator": "yoho_b200.estimator",
}


_METRIC_FUNCS = ("rotation_error", "translation_error", "computeTransformationErr", "evaluate_registration")


def install(metrics=False):
    for ref_name, ours in _ALIASES.items():
        sys.modules[ref_name] = importlib.import_module(ours)
    done = sorted(_ALIASES)
    if metrics:
        import types
        for name in ("nibabel", "nibabel.quaternions"):     # utils/RR_cal.py:10 imports it at module level
            try:
                importlib.import_module(name)
            except ImportError:
                sys.modules[name] = types.ModuleType(name)
        ours = importlib.import_module("yoho_b200.rr_cal")
        ref = importlib.import_module("utils.RR_cal")       # the reference's module (its root must be on sys.path)
        for fn in _METRIC_FUNCS:
            setattr(ref, fn, getattr(ours, fn))
        done.append("utils.RR_cal:" + ",".join(_METRIC_FUNCS))
    return done
