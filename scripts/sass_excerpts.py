"""Writes profiles/r02_sass_<kernel>.txt: mnemonic counts and the tensor-core / TMA / mbarrier instruction lines (with two lines
of context) of the hot kernels, from `cuobjdump -sass` of the built library.  The .so itself is git-ignored; these excerpts
are the committed evidence that the kernels are tcgen05 / TMA code.  Usage: python scripts/sass_excerpts.py"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "csm_mlx_b200", "libcsm_b200.so")
KERNELS = {"k_gemm_part_t<false, false>": "gemm_part", "k_gemm_part_t<true, false>": "gemm_part_swiglu",
           "k_gemm_part_t<true, true>": "gemm_part_swiglu_cta_pair", "k_linear_tc": "linear_tc",
           "fk_bf16::k_frame(": "frame", "fk_e4m3::k_frame(": "frame_e4m3", "k_gemm_tc3": "gemm_tc3",
           "k_attn_decode_small<128>": "attn_decode_small", "k_attn_decode_chunked<64>": "attn_decode_chunked",
           "k_resid_norm_split<1>": "resid_norm_split"}
KEY = re.compile(r"UTCHMMA|UTCBAR|UTMALDG|UCGABAR|UBLKCP|UBLKPF|LDTM|UTCATOMSWS|SYNCS\.(ARRIVE|EXCH|PHASECHK)|ACQBULK|PREEXIT|UTMAPF|FENCE\.VIEW\.ASYNC|F2FP\.BF16|LDGSTS|LDGDEPBAR|DEPBAR|F2FP\.F16\.E4M3|F2F")

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
blocks = re.split(r"\n\s*Function : ", sass)
for blk in blocks[1:]:
    name_m, _, body = blk.partition("\n")
    name = subprocess.run(["c++filt", name_m.strip()], capture_output=True, text=True).stdout.strip()
    tag = next((v for k, v in KERNELS.items() if ("csmb::" + k) in name), None)
    if tag is None:
        continue
    lines = [l for l in body.splitlines() if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", l) and not re.match(r"\s*/\* 0x", l)]
    ops = collections.Counter()
    for l in lines:
        m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if m:
            ops[".".join(m.group(1).split(".")[:3])] += 1
    out = [f"# {name}", f"# cuobjdump -sass csm_mlx_b200/libcsm_b200.so (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a); {len(lines)} instructions",
           "# counts of tensor-core / TMA / TMEM / mbarrier / PDL instructions:"]
    for op, n in sorted(ops.items()):
        if KEY.search(op):
            out.append(f"#   {op:40s} {n}")
    out.append("# excerpt: every such instruction with two lines of context")
    keep = set()
    for i, l in enumerate(lines):
        if KEY.search(l):
            keep.update(range(max(0, i - 2), min(len(lines), i + 3)))
    prev = -2
    shown = 0
    for i in sorted(keep):
        if shown > 400:
            out.append("        ...")
            break
        if i != prev + 1:
            out.append("        ...")
        out.append(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", lines[i]).rstrip())
        prev = i
        shown += 1
    with open(os.path.join(ROOT, "profiles", f"r02_sass_{tag}.txt"), "w") as f:
        f.write("\n".join(out) + "\n")
    print(tag, len(lines), {k: v for k, v in ops.items() if KEY.search(k)})
