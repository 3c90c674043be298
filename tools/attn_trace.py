"""Print the per-phase cycle breakdown written by WB_ATTN_TRACE (see wb_dbg_attention)."""
import sys
v = [int(x) for x in open(sys.argv[1]).read().split()]
t0 = min(x for x in v if x > 0)
print("softmax threads (cycles rel. to start):  j wg | wait_start  s_ready  ld_done  max_done  exp_done  arrived | wait  ld  max  exp  arrive")
for j in range(12):
    for wg in range(2):
        b = (j * 2 + wg) * 8
        a = [v[b + k] - t0 for k in range(6)]
        print(f"{j:2d} {wg} | " + " ".join(f"{x:7d}" for x in a) + " | " + " ".join(f"{a[k+1]-a[k]:6d}" for k in range(5)))
print("control thread: j wg | wait_p_start  p_ready  v_ready  committed | wait_p  wait_v  issue")
for j in range(12):
    for wg in range(2):
        b = 512 + (j * 2 + wg) * 4
        a = [v[b + k] - t0 for k in range(4)]
        print(f"{j:2d} {wg} | " + " ".join(f"{x:7d}" for x in a) + " | " + " ".join(f"{a[k+1]-a[k]:6d}" for k in range(3)))
