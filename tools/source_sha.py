"""sha256 (first 16 hex digits) of every kernel source of the product library, as JSON on stdout.  tools/round_profile.sh runs it on
the GPU box next to the ncu captures, so that profiles/traffic.json can say which sources its numbers were measured on; bench.py
recomputes the hashes and refuses to report `roofline.traffic` from a capture of different sources."""
import hashlib
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def source_sha():
    out = {}
    for d in ("alphazero_risk_b200/csrc", "include"):
        for f in sorted(os.listdir(os.path.join(ROOT, d))):
            if f.endswith((".cu", ".cuh", ".h", ".cpp")):
                out[f] = hashlib.sha256(open(os.path.join(ROOT, d, f), "rb").read()).hexdigest()[:16]
    return out


if __name__ == "__main__":
    print(json.dumps(source_sha(), indent=1))
