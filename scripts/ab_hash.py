"""Developer tool: SHA-256 of the stepper's output on the check_variant workload for the library selected by
SCRUTINY_B200_LIB, so that A/B builds (scripts/ab_build.sh) can be compared bit for bit across processes.
usage: SCRUTINY_B200_LIB=scrutiny_b200/_ab/NAME.so python scripts/ab_hash.py [net ...]"""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import check_variant  # noqa: E402

for net in (sys.argv[1:] or ["powerlaw", "trrust"]):
    out = check_variant.run(net, {})
    print(net, hashlib.sha256(out.tobytes()).hexdigest()[:16], flush=True)
