"""crb / extract golden cases (tests/golden/tags, recorded from the unmodified reference by scripts/make_golden_tags.py) run through
fastf_b200.tags_host; shared by the emulator (CPU) and the GPU tests."""
import gzip
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = os.path.join(ROOT, "tests", "golden", "tags")


def cases():
    return json.load(open(os.path.join(D, "manifest.json")))["cases"]


def run_case(ctx, case, tmp):
    from fastf_b200 import tags_host as T
    bam = os.path.join(D, case["input"])
    want = gzip.open(os.path.join(D, case["expect"]), "rb").read()
    if case["kind"] == "crb":
        out = os.path.join(tmp, case["name"] + ".gz")
        assert T.crb(ctx, bam, out) == case["reads"]
        got = gzip.open(out, "rb").read()
    else:
        total, valid = T.extract_bam(ctx, bam, case["tag"], case["type"], tmp)
        assert (total, valid) == (case["total"], case["valid"]), case["name"]
        got = open(os.path.join(tmp, "tag_summary.csv"), "rb").read()
    assert got == want, case["name"]


if __name__ == "__main__":   # emulator entry: FASTF_GPU_LIB points at the SIMT-emulator build
    import sys
    import tempfile
    sys.path.insert(0, ROOT)
    from fastf_b200 import _lib
    ctx = _lib.Context(0)
    with tempfile.TemporaryDirectory() as tmp:
        for c in cases():
            if len(sys.argv) > 1 and c["name"] not in sys.argv[1:]:
                continue
            run_case(ctx, c, tmp)
            print("ok", c["name"])
        # streaming: three blocks per chunk (cross-chunk merge of counts and first occurrences); forced hash collisions in round 0
        ctx.lib.fastf_taghist_test_hooks(ctx.h, 3, 0xff)
        for c in cases():
            if c["input"] == "tags.bam":
                run_case(ctx, c, tmp)
                print("ok-streamed-collided", c["name"])
        from fastf_b200 import tags_host as T
        import numpy as np
        st, _ = T.taghist(ctx, np.fromfile(os.path.join(D, "tags.bam"), dtype=np.uint8), "CB", 0, "CR")
        assert st["hash_rounds"] == 2, st
        ctx.lib.fastf_taghist_test_hooks(ctx.h, 0, 0)
