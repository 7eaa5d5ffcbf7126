"""Per-tree-level view of an ncu launch list (gpu__time_duration.sum CSV) of `tools/prof_one.py N 2`:
the LAST solve of the list, cut at the first kernel of every join (k_hash_insert).  Segment L = join + solve
of level L followed by the Transform batch that re-bases its outputs / prepares the End maps of level L+1."""
import csv, collections, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; data = rows[hdr + 1:]
kn, mv, mu = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
L = []
for r in data:
    name = re.sub(r'\(.*', '', r[kn]).replace('<unnamed>::', '').replace('void ', '')
    name = re.sub(r'<.*', '', name)
    L.append((name, float(r[mv].replace(',', '')) * {'us': 1e-3, 'ns': 1e-6, 's': 1e3, 'ms': 1.0}.get(r[mu], 1.0)))
joins = [i for i, (n, _) in enumerate(L) if n == 'k_hash_insert']
nlev = int(sys.argv[2]) if len(sys.argv) > 2 else 12
joins = joins[-nlev:]
# the level-0 Transform of the last solve starts after the previous solve's last k_pp_reduce / k_tf_posefin
prev_end = max(i for i, (n, _) in enumerate(L[:joins[0]]) if n == 'k_backsub') if any(n == 'k_backsub' for n, _ in L[:joins[0]]) else -1
first = prev_end + 1
# skip the previous solve's final re-base (one Transform batch) if there is one: the batch of level 0 is the LAST k_find_pos before joins[0]
fp = [i for i in range(first, joins[0]) if L[i][0] == 'k_find_pos']
start0 = fp[-1] if fp else first
big = ['tfc::k_tf_chunk', 'schur_lock::k_schur_lock', 'schur_pipe::k_schur_pipe', 'k_join_w', 'k_front_factor', 'k_backsub', 'k_pat_chunk']
cols = ['k_tf_chunk', 'Schur', 'k_join_w', 'k_front_factor', 'k_backsub', 'k_pat_chunk']
print("| segment | launches | kernel time ms | " + " | ".join(cols) + " | the other launches (count: ms) |")
print("|---|---:|---:|" + "---:|" * (len(cols) + 1))
def row(label, seg):
    agg = collections.defaultdict(float)
    for n, v in seg: agg[n] += v
    tot = sum(v for _, v in seg)
    vals = [agg['tfc::k_tf_chunk'], agg['schur_lock::k_schur_lock'] + agg['schur_pipe::k_schur_pipe'], agg['k_join_w'],
            agg['k_front_factor'], agg['k_backsub'], agg['k_pat_chunk']]
    others = [(n, v) for n, v in seg if n not in big]
    print(f"| {label} | {len(seg)} | {tot:.3f} | " + " | ".join(f"{v:.3f}" for v in vals) +
          f" | {len(others)}: {sum(v for _, v in others):.3f} |")
    return tot
tot = row("Transform of the level-0 End maps", L[start0:joins[0]])
bounds = joins + [len(L)]
for li in range(nlev):
    tot += row(f"level {li}", L[bounds[li]:bounds[li + 1]])
print(f"\nlast solve: {tot:.2f} ms of kernel time over {len(L) - start0} launches (cold-cache, serialised: compare shares)")
