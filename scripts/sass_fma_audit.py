#!/usr/bin/env python
"""Where do FFMA / DFMA come from in a library compiled with -fmad=false?

The reference object code has no fused multiply-adds, and every voxel index / comparison in the kernels must round like it
(DESIGN.md section 4).  -fmad=false stops nvcc from CONTRACTING a*b+c; the FMAs that remain are the ones inside the
IEEE-correct software sequences for float / double division, reciprocal and square root (MUFU.RCP / MUFU.RSQ seed + Newton
steps), whose results are correctly rounded whatever they use internally.  This script checks that claim on the SASS:
for every kernel it counts FFMA / DFMA and how many of them sit within `WINDOW` instructions after a MUFU seed, or are one of
the sequences' recognisable slow-path steps (exact scaling by a power of two with a zero addend, the directed-rounding .RM / .RP
/ .RZ step that ends a correctly-rounded double sqrt / div, the +INF special-value path of sqrtf); anything else is listed.

    python scripts/sass_fma_audit.py [path/to/libgoicp_b200.so] > profiles/r2_sass_fma_audit.txt
"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "cuda-go-icp_b200", "libgoicp_b200.so")
WINDOW = 96


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout.splitlines()
    fn, ops = None, collections.OrderedDict()
    for l in sass:
        m = re.search(r"Function : (\S+)", l)
        if m:
            fn = m.group(1); ops[fn] = []; continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);", l)
        if m and fn:
            ops[fn].append((m.group(1), m.group(2)))
    names = subprocess.run(["c++filt"], input="\n".join(ops), capture_output=True, text=True).stdout.splitlines()
    print(f"# {os.path.basename(LIB)}: FFMA / DFMA per kernel; 'explained' = within {WINDOW} instructions after a MUFU.RCP/RSQ(64H) seed,")
    print("# i.e. inside a correctly-rounded division / reciprocal / square-root sequence (nvcc -fmad=false -prec-div=true -prec-sqrt=true)")
    print("kernel,FFMA,DFMA,explained,unexplained")
    tot = [0, 0, 0, 0]
    for (f, seq), name in zip(ops.items(), names):
        last_seed = -10 ** 9
        n_f = n_d = expl = 0
        odd = []
        for i, (op, rest) in enumerate(seq):
            base = op.split(".")[0]
            if base == "MUFU" and ("RCP" in op or "RSQ" in op or "SQRT" in op):
                last_seed = i
            if base in ("FFMA", "DFMA"):
                n_f += base == "FFMA"; n_d += base == "DFMA"
                exact_scaling = rest.rstrip().endswith(", RZ") and re.search(r", [0-9.e+-]+, RZ\s*$", rest) is not None     # x * 2^k + 0 (denormal handling of the slow paths)
                directed = ".RM" in op or ".RP" in op or ".RZ" in op                                                    # last step of the correctly-rounded double sqrt / div
                special = "INF" in rest                                                                                # special-value path of sqrtf
                if i - last_seed <= WINDOW or exact_scaling or directed or special:
                    expl += 1
                else:
                    odd.append(f"{i}:{op}{rest}")
        if n_f or n_d:
            short = re.sub(r"\(.*$", "", name.replace("(anonymous namespace)::", "")).replace("void ", "")
            print(f"\"{short}\",{n_f},{n_d},{expl},{n_f + n_d - expl}")
            for o in odd[:6]:
                print(f"#    unexplained: {o}")
            tot[0] += n_f; tot[1] += n_d; tot[2] += expl; tot[3] += n_f + n_d - expl
    print(f"# total: FFMA {tot[0]}, DFMA {tot[1]}, explained {tot[2]}, unexplained {tot[3]}")


if __name__ == "__main__":
    main()
