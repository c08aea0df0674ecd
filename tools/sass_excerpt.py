"""Static SASS mnemonic counts of the shipped library (no GPU needed): library totals and the hot kernels.
python tools/sass_excerpt.py [lattigo_b200/lib/liblattigo_b200.so] > profiles/rNN_sass_excerpt.txt
UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier, LDG.*256 = 256-bit global loads, SHFL = warp shuffles."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOT = ("ks_strided_j4_kernel<4, 4, true, true>", "fz_chunk_epi_fp8_kernel", "ks_chunk_mac_int8r_kernel<false>", "ks_chunk_mac_fp8r_kernel<1>",
       "ks_chunk_mac_fp8r_kernel<0>", "lt_hoisted_auto_kernel", "ntt_persist2_kernel<lgpu::IntFwdOps2<4, false> >", "ntt_persist2_kernel<lgpu::FpFwdOps2<4> >",
       "ntt_persist_tma_kernel<4>", "ntt_persist3_kernel<lgpu::FpFwdOps3<4> >", "ks_prepare_kernel", "ckks_tensor_kernel")
KEYS = ("UBLKCP", "SYNCS", "LDGSTS", "SHFL", "DFMA", "DADD", "DMUL", "IMAD", "LDG", "LDG256", "STG", "LDS", "STS", "LDL", "STL", "BAR", "UTMALDG")


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "lattigo_b200", "lib", "liblattigo_b200.so")
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    filt = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.splitlines()
    names = iter(filt)
    per = collections.OrderedDict()
    cur = None
    for ln in out.splitlines():
        if "Function :" in ln:
            cur = per.setdefault(next(names), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", ln)
        if m and cur is not None:
            op = m.group(1)
            cur["total"] += 1
            base = op.split(".")[0]
            cur[base] += 1
            if base == "LDG" and ".256" in op:
                cur["LDG256"] += 1
    tot = collections.Counter()
    for c in per.values():
        tot.update(c)
    print("cuobjdump -sass %s (sm_100a), %d kernels; mnemonic counts (static instructions; tools/sass_excerpt.py)" % (os.path.relpath(so, ROOT), len(per)))
    print("library totals: " + ", ".join("%s=%d" % (k, tot[k]) for k in KEYS))
    for h in HOT:
        for name, c in per.items():
            if h in name:
                print(name)
                print("    total=%d  " % c["total"] + "  ".join("%s=%d" % (k, c[k]) for k in KEYS if c[k]))


if __name__ == "__main__":
    main()
