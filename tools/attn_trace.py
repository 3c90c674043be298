"""Print the per-phase cycle breakdown written by WB_ATTN_TRACE (see wb_dbg_attention)."""
import sys
v = [int(x) for x in open(sys.argv[1]).read().split()]
NWG = int(sys.argv[2]) if len(sys.argv) > 2 else 1   # softmax warpgroups per CTA in the traced build
t0 = min(x for x in v if x > 0)
print("softmax thread (row 0), cycles rel. to start:  j wg | wait_start  s_ready  max_done  xchg_done  arrived | wait  max  xchg  exp")
for j in range(24):
    for wg in range(NWG):
        b = (j * 2 + wg) * 8
        a = [v[b + k] - t0 for k in range(5)]
        print(f"{j:2d} {wg} | " + " ".join(f"{x:7d}" for x in a) + " | " + " ".join(f"{a[k+1]-a[k]:6d}" for k in range(4)))
print("MMA warp: j wg | wait_p0_start  p0_ready  p1_ready  committed | wait_p0  pv0+wait_p1  pv1+s")
for j in range(24):
    for wg in range(NWG):
        b = 512 + (j * 2 + wg) * 4
        a = [v[b + k] - t0 for k in range(4)]
        print(f"{j:2d} {wg} | " + " ".join(f"{x:7d}" for x in a) + " | " + " ".join(f"{a[k+1]-a[k]:6d}" for k in range(3)))
