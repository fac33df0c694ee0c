"""profiles/sass_r2.txt: static SASS instruction census per kernel of libphare_b200.so (cuobjdump -sass), the mnemonics that
prove what each kernel is built from (UBLKCP = cp.async.bulk / 1-D TMA, SYNCS = mbarrier, LDGSTS = cp.async, LDS, DFMA ...).
usage: python tools/sass_census.py [out]"""
import collections
import re
import subprocess
import sys

LIB = "phare_b200/lib/libphare_b200.so"
WANT = ['tile_kernel', 'push_tma_kernel', 'push_plan_kernel', 'deposit_scatter_kernel', 'deposit_cells_kernel', 'bin_count_kernel',
        'bin_scatter_kernel', 'box_op_batch', 'faraday_kernel', 'ampere_kernel', 'ohm_kernel', 'push_deposit_cells_kernel',
        'peer_signal', 'peer_wait', 'gather_kernel']
MN = ['UBLKCP', 'SYNCS', 'LDGSTS', 'LDS', 'STS', 'LDG', 'STG', 'DFMA', 'DMUL', 'DADD', 'REDG.E.ADD.F64', 'ATOMG', 'REDUX', 'MATCH',
      'SHFL', 'MUFU.RCP64H', 'BAR.SYNC']
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
rows = []
for f in re.split(r'\n\s*Function : ', txt)[1:]:
    name = f.split('\n', 1)[0].strip()
    dem = subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip()
    if not any(w in dem for w in WANT):
        continue
    cnt, ninst = collections.Counter(), 0
    for line in f.split('\n'):
        m = re.search(r'/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if not m:
            continue
        ninst += 1
        op = m.group(1)
        for k in MN:
            if op == k or op.startswith(k + '.'):
                cnt[k] += 1
    short = re.sub(r'\(int\)|\(bool\)', '', re.sub(r'\(phb::.*', '', dem))[:100]
    rows.append((short, ninst, cnt))
rows.sort()
out = ["# SASS census of libphare_b200.so (cuobjdump -sass, sm_100a): static instruction counts per kernel instantiation",
       "# UBLKCP = cp.async.bulk (1-D TMA bulk copy), SYNCS = mbarrier ops, LDGSTS = cp.async, REDG.E.ADD.F64 = FP64 global atomics",
       "# tile_kernel<DIM, ORDER, GS, EXACT, DEPOSIT, WRITE, PLAN>; push_tma_kernel<DIM, ORDER, EXACT, HAS_FIRST, COPY_WQ, PLAN>",
       "# kernel | instructions | " + " ".join(MN)]
for short, ninst, cnt in rows:
    out.append(f"{short} | {ninst} | " + " ".join(str(cnt[k]) for k in MN))
open(sys.argv[1] if len(sys.argv) > 1 else "profiles/sass_r2.txt", "w").write("\n".join(out) + "\n")
print(len(rows), "kernels")
