"""Make the reference's own entry scripts (`Test.py`, `Demo.py`) run on the B200 backend UNCHANGED.

    import yoho_b200.dropin; yoho_b200.dropin.install()     # before `import tests.evaluator`

`install()` registers this package's modules under the names the reference imports —
`utils.network`, `utils.knn_search`, `tests.extractor`, `tests.matcher`, `tests.estimator` — so
`tests/evaluator.py:22-24` and `Demo.py:6-12` pick up the B200 implementations through the reference's own
`name2*` registries.  Everything else (`parses`, `utils.dataset`, `utils.RR_cal`, `tests.evaluator`) stays the
reference's.  See INTEGRATION.md.
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


def install():
    for ref_name, ours in _ALIASES.items():
        sys.modules[ref_name] = importlib.import_module(ours)
    return sorted(_ALIASES)
