"""Writes profiles/sass_k_scene_mix.txt: what the shipped callback kernel compiles to (sm_100a), as evidence that it is
hand-written Blackwell code - mnemonic counts (UBLKCP = cp.async.bulk / TMA bulk copy, SYNCS = mbarrier, FADD2 / FFMA2 =
packed FP32x2, no HMMA / UTC*MMA: this path has no dense contraction) and the inner loop of the common case.

    python tools/sass_excerpt.py
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oddio_b200", "liboddio_b200.so")
OUT = os.path.join(ROOT, "profiles", "sass_k_scene_mix.txt")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)
    pick = [f for f in funcs if f.startswith("_ZN4odbk11k_scene_mixINS_6SmxCfgILi2ELi16ELi0ELi4EEELb0ELb0EEE")]
    assert pick, "the default FMA instantiation of k_scene_mix is not in the library"
    body = pick[0]
    ins = re.findall(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", body)
    counts = collections.Counter(i.split(".")[0] for i in ins)
    lines = body.splitlines()
    # the common case's inner loop: from the first LDS.64 checkpoint load after the first FADD2.RM back to the FFMA2 accumulate
    idx = [i for i, l in enumerate(lines) if "FADD2.RM" in l]
    start = max(0, idx[0] - 14) if idx else 0
    excerpt = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l).rstrip() for l in lines[start:start + 96] if "/*" in l and not re.match(r"\s*/\* 0x", l)]
    with open(OUT, "w") as f:
        f.write("k_scene_mix<SmxCfg<2,16,0,4>, STRICT=false, VARBATCH=false> in oddio_b200/liboddio_b200.so (cuobjdump -sass), sm_100a\n")
        f.write(f"{len(ins)} instructions. Mnemonic counts (top 28):\n")
        for k, v in counts.most_common(28):
            f.write(f"  {k:10s} {v}\n")
        for k in ("UBLKCP", "SYNCS", "FADD2", "FFMA2", "FMUL2", "HMMA", "UTCHMMA", "UTMALDG", "ELECT", "R2UR"):
            f.write(f"{k}: {counts.get(k, 0)}\n")
        f.write("\nInner loop of the common case (full tile, both ears on the doppler path), 4 frames x 2 ears in flight:\n")
        f.write("\n".join(excerpt) + "\n")
    print(open(OUT).read()[:1500])


if __name__ == "__main__":
    main()
