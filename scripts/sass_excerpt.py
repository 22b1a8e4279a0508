#!/usr/bin/env python
"""profiles/r02_raster_sass_hot_loop.txt: SASS of the rasterizer's hot paths, cut out of `cuobjdump -sass` of the built
library (development aid; run after geograypher_b200/csrc/build.sh)."""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
so = ROOT / "geograypher_b200" / "libgeograypher_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True, check=True).stdout


def function(tag):
    out, on = [], False
    for line in txt.splitlines():
        if "Function :" in line:
            on = tag in line
            continue
        if on:
            m = re.match(r"\s*/\*([0-9a-f]{4})\*/\s+(.*?);", line)
            if m:
                out.append((int(m.group(1), 16), m.group(2).strip()))
    return out


def mix(ins, top=14):
    c = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", s).split()[0].split(".")[0] for _, s in ins)
    return ", ".join(f"{k} {v}" for k, v in c.most_common(top))


def dump(ins):
    return "\n".join(f"  /*{a:04x}*/  {s}" for a, s in ins)


k4 = function("k_raster_tilesILi4EfLi0")
kd = function("k_raster_tilesILi2EfLi10")
# staging loop: from the first LDG.E.128.CONSTANT back to the address set-up, up to the loop's BRA
i_ldg = next(i for i, (_, s) in enumerate(k4) if "LDG.E.128.CONSTANT" in s)
i_end = next(i for i in range(i_ldg, len(k4)) if k4[i][1].startswith("@!P0 BRA") or k4[i][1].startswith("@P0 BRA"))
staging = k4[max(0, i_ldg - 8): i_end + 2]
# per-face fast path: from the lane-mask LDS.128 [..+0x30] to the BRA that closes the 8-pixel block
i_face = next(i for i, (_, s) in enumerate(k4) if "LDS.128" in s and "+0x30]" in s)
i_face_end = next(i for i in range(i_face, len(k4)) if k4[i][1].startswith("BRA "))
face = k4[i_face: i_face_end + 1]
# winners epilogue (interior tile): the block that holds the LAST eight REDG of the kernel
reds = [i for i, (_, s) in enumerate(k4) if s.startswith("REDG") or " REDG" in s]
epi = k4[reds[-8] - 34: reds[-1] + 2] if len(reds) >= 8 else []
i_pf = next((i for i, (_, s) in enumerate(kd) if "UBLKPF" in s), None)
pf = kd[i_pf - 9: i_pf + 3] if i_pf is not None else []

out = [
    "# cuobjdump -sass geograypher_b200/libgeograypher_b200.so (sm_100a), round 2 final build (scripts/sass_excerpt.py)",
    f"# k_raster_tiles<GG_RM_WINNERS_ONLY, float, 0>: {len(k4)} SASS instructions in total",
    f"#   mix: {mix(k4)}",
    "#",
    "# 1. staging copy of a tile's 64-byte setups into shared memory (rolled loop: one pass for <= 8 faces)",
    dump(staging),
    "#",
    "# 2. per-face fast path: lane-mask test, tile-relative edge values and 1/z of the lane's 8 pixels, then per pixel",
    "#    FFMA (1/z) + 3 IMAD.IADD (edge stepping) + LOP3 (e0|e1|e2) + ISETP (coverage, folded with the depth predicate)",
    "#    + DSETP.GT (64-bit key = 1/z bits : ~face, compared as ONE float64 on the FP64 pipe) + 2 FSEL (key pair);",
    "#    IMAD.MOV = register-pair assembly.  No list position is carried: the key's low word IS the face.",
    f"#   mix of this block ({len(face)} instructions per face): {mix(face)}",
    dump(face),
    "#",
    "# 3. winners epilogue of an interior tile: one max-reduction per run-end straight into the view's winner array",
    "#    (ptxas turns the predicated red.global into a short branch; the address is wbase - 4 * (int)nf)",
    f"#   mix of this block ({len(epi)} instructions): {mix(epi)}",
    dump(epi),
    "#",
    f"# 4. k_raster_tiles<GG_RM_DENSE, float, 10>: {len(kd)} SASS instructions; the bulk L2 prefetch of the tile's score rows",
    "#    (cp.async.bulk.prefetch.L2 -> UBLKPF.L2; ptxas serialises the 8 issuing lanes through the uniform datapath)",
    dump(pf),
    "# No UTMALDG / UTMASTG: the rasterizer is ALU / issue bound and streams 64-byte setups, and the dense epilogue consumes",
    "# its scores straight from L2 with LDG.64 after the bulk prefetch (a shared-memory staging of the score tile costs the",
    "# occupancy the rasterizer phase needs: DESIGN.md section 5).",
]
dst = ROOT / "profiles" / "r02_raster_sass_hot_loop.txt"
dst.write_text("\n".join(out) + "\n")
print(f"wrote {dst}: staging {len(staging)}, face {len(face)}, epilogue {len(epi)}, prefetch {len(pf)} instructions")
