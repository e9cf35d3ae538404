"""TEST INFRASTRUCTURE ONLY -- records the public function surface of the unmodified reference module
(RNAscrutiny/RNAscrutiny/MatriSeq.py: names, parameter names, defaults) as tests/golden/matriseq_signatures.json, the
fixture tests/test_host_surface.py::test_module_surface_matches_the_reference checks scrutiny_b200.MatriSeq against.
Run in the build container only (needs /root/reference): python oracle/make_signatures.py"""
import inspect
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import reference_harness  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "matriseq_signatures.json")


def main():
    ms = reference_harness.load_reference()
    table = {}
    for name, fn in sorted(vars(ms).items()):
        if not inspect.isfunction(fn) or fn.__module__ != ms.__name__ or name.startswith("_"):
            continue
        params = []
        for p in inspect.signature(fn).parameters.values():
            params.append([p.name, None if p.default is inspect.Parameter.empty else repr(p.default)])
        table[name] = {"line": fn.__code__.co_firstlineno, "params": params}
    with open(OUT, "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)
    print("wrote", OUT, len(table), "functions")


if __name__ == "__main__":
    main()
