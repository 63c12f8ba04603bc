"""Golden vectors for the post-composite volume smoothing: runs the reference's own code --
generate.py:254-263 (exec of the cited lines) calling modules/postprocess.py:postprocess_ct_volume unmodified (it only
needs numpy + scipy, both present) -- on small seeded volumes and stores input seeds + int16 outputs in
tests/golden/postprocess.npz.   usage: python oracle/make_golden_postprocess.py   (in the container that has /root/reference)
TEST INFRASTRUCTURE ONLY."""
import os
import sys
import textwrap

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import ducosy_oracle as orc  # noqa: E402

sys.path.insert(0, REF)
from modules.postprocess import postprocess_ct_volume  # noqa: E402  (the reference's own function)


def ref_lines(relpath, first, last):
    with open(os.path.join(REF, relpath)) as f:
        lines = f.readlines()
    return textwrap.dedent("".join(lines[first - 1:last]))


def main():
    src = ref_lines("generate.py", 254, 263)
    out = {}
    for name, (S, H, W, seed) in {"a": (12, 48, 64, 5), "b": (2, 33, 40, 6), "c": (7, 64, 64, 7)}.items():
        vol = orc.postprocess_test_volume(S, H, W, seed)
        env = {"np": np, "postprocess_ct_volume": postprocess_ct_volume, "merged_volume": [v for v in vol]}
        exec(src, env)
        res = env["merged_volume"]
        assert res.dtype == np.int16 and res.shape == vol.shape
        assert np.array_equal(res, orc.postprocess_volume(vol)), name          # the oracle restatement is pinned here
        out[f"shape_{name}"] = np.array([S, H, W, seed])
        out[f"out_{name}"] = res
    np.savez_compressed(os.path.join(OUT, "postprocess.npz"), **out)
    print("postprocess golden ok", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
