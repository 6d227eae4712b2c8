"""tools_trace.py trace.bin — timeline of CTA 0's pipeline from a PIPE_TRACE build (SM clocks; 1.965 GHz)."""
import sys, numpy as np, collections
d = np.fromfile(sys.argv[1], np.uint64).reshape(-1, 2)
t = d[:, 0].astype(np.int64); w = d[:, 1]
code = (w >> np.uint64(56)).astype(int); warp = ((w >> np.uint64(48)) & np.uint64(0xff)).astype(int)
a = ((w >> np.uint64(24)) & np.uint64(0xffffff)).astype(int); b = (w & np.uint64(0xffffff)).astype(int)
t0 = t.min(); us = (t - t0) / 1965.0
print("events", len(t), "span %.1f us" % us.max())
order = np.argsort(t, kind="stable")
# producer: per tile
P = collections.defaultdict(dict)
for i in order:
    if code[i] <= 5: P[a[i]][code[i]] = (us[i], warp[i], b[i])
print("tile  warp  total |  start  empty_ok  alloc_ok  issued  arrived   (us)")
for k in sorted(P):
    e = P[k]
    print("%4d  %4d  %5d | " % (k, e[1][1], e.get(3, (0, 0, 0))[2]) + "  ".join("%7.2f" % e[c][0] if c in e else "   -   " for c in (1, 2, 3, 4, 5)))
# consumers: time per state
st = collections.Counter(); cnt = collections.Counter()
names = {(17, 18): "wait payload (full)", (18, 14): "own fields", (13, 17): "claim+seek+prefetch", (10, 11): "wait table", (12, 13): "cold fetch", (13, 14): "seek+prefetch setup", (14, 15): "pair loop", (15, 16): "epilogue", (16, 12): "to next batch", (11, 12): "enter->batch", (16, 10): "batch end->enter", (11, 10): "enter->enter", (11,13):"?"}
last = {}
full_at = {}
for i in order:
    c = code[i]
    if c < 10: continue
    wv = warp[i]
    if c == 11: full_at.setdefault(a[i], us[i])
    if wv in last:
        lc, lt = last[wv]
        st[(lc, c)] += us[i] - lt; cnt[(lc, c)] += 1
    last[wv] = (c, us[i])
tot = sum(st.values())
print("consumer warps: time by state (sum over warps, us) of %.1f total" % tot)
for kx, v in sorted(st.items(), key=lambda kv: -kv[1]):
    print("  %-22s %9.1f us  %5.1f%%  n=%d  mean %.2f us" % (names.get(kx, str(kx)), v, 100 * v / tot, cnt[kx], v / cnt[kx]))
print("first time a consumer saw tile k full vs producer arrive:")
for k in sorted(P):
    if k in full_at and 5 in P[k]: print("  tile %d: producer arrived %.2f, first consumer through %.2f" % (k, P[k][5][0], full_at[k]))
# batch durations
# per tile: consumption window
first = {}; lastend = {}; nb = collections.Counter()
for i in order:
    if code[i] == 12: first.setdefault(a[i], us[i]); nb[a[i]] += 1
    if code[i] == 16: lastend[a[i]] = us[i]
print("tile: batches, first batch start, last batch end, (ring alloc of tile), full")
for k in sorted(first):
    print("  %3d: %3d  %7.2f  %7.2f   alloc %7.2f  full %7.2f" % (k, nb[k], first[k], lastend.get(k, -1), P.get(k, {}).get(4, (0,))[0], P.get(k, {}).get(5, (0,))[0]))
# per warp timeline of batch starts
bw = collections.defaultdict(list)
for i in order:
    if code[i] == 12: bw[warp[i]].append((us[i], a[i], b[i] >> 8))
for wv in sorted(bw)[:8]:
    print("warp", wv, " ".join("%.1f:k%d.b%d" % x for x in bw[wv]))
