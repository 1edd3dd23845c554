#!/usr/bin/env python
"""Build an A/B variant of the library next to the product build.

    python scripts/build_variant.py NAME [-DMACRO=VALUE ...]     ->  gpurun_variants/lib_NAME.so

Same sources and flags as cap4d_b200/build.py plus the given -D switches (the opt-in experiments of the kernels, e.g.
-DCAP4D_ATTN_PACKED_F32X2=1, -DCAP4D_ATTN_REGS_CTRL=32 -DCAP4D_ATTN_REGS_SOFTMAX=112).  The product library is not
touched; scripts/attn_variants.sh / gemm_variants.sh swap the variants in on the GPU box and run tests + timings.
"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cap4d_b200 import build as B  # noqa: E402


def main():
    if len(sys.argv) < 2 or sys.argv[1].startswith("-"):
        sys.exit(__doc__)
    name, defs = sys.argv[1], sys.argv[2:]
    out_dir = os.path.join(ROOT, "gpurun_variants")
    os.makedirs(out_dir, exist_ok=True)
    lib = os.path.join(out_dir, f"lib_{name}.so")
    nvcc = B._nvcc()
    with tempfile.TemporaryDirectory() as tmp:
        procs, objs = [], []
        for src in B.SOURCES:
            obj = os.path.join(tmp, src.replace(".cu", ".o"))
            objs.append(obj)
            cmd = [nvcc] + B.NVCC_FLAGS + B.EXTRA_FLAGS.get(src, []) + defs + ["-c", os.path.join(B.CSRC, src), "-o", obj]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        for src, p in procs:
            out, _ = p.communicate()
            if p.returncode != 0:
                sys.stdout.write(out.decode(errors="replace"))
                sys.exit(f"nvcc failed on {src}")
        subprocess.check_call([nvcc, "-shared", "-cudart", "shared", "-o", lib] + objs)
    print(lib)


if __name__ == "__main__":
    main()
