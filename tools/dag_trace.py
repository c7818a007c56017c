"""Dev tool: read the per-ticket time stamps k_dag2 writes with PB200_DAG_TRACE=<file> (engine.cu, solve_tf) and say
where a sweep spends its time.  usage: python tools/dag_trace.py trace.bin
record = {taken, dependencies met, done, (sm << 32) | sub-tiles << 16 | stages in flight << 8 | is-diagonal, copies landed,
input vector in shared memory, partial sums written, before the fence}, ns of %globaltimer."""
import sys

import numpy as np


def main():
    tr = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(2, -1, 8)
    G = tr.shape[1]
    for d, name in enumerate(("down", "up")):
        t = tr[d]
        ok = t[:, 2] > 0
        take, dep, end = (t[ok, k].astype(np.int64) for k in range(3))
        isd = (t[ok, 3] & 1).astype(bool)
        nq = ((t[ok, 3] >> 8) & 0xff).astype(int)
        sm = (t[ok, 3] >> 32).astype(int)
        t0 = take.min()
        span = end.max() - t0
        print(f"== {name}: {ok.sum()} / {G} tickets, span {span / 1e3:.1f} us, timer step {np.min(np.diff(np.unique(end))) if len(end) > 1 else 0} ns, "
              f"{len(np.unique(sm))} SMs")
        for lab, m in (("T", ~isd), ("D", isd)):
            if not m.any():
                continue
            w = (dep - take)[m]; p = (end - dep)[m]
            ph = [np.median((t[ok, b].astype(np.int64) - t[ok, a].astype(np.int64))[m]) / 1e3 for a, b in ((0, 2), (1, 5), (5, 7), (7, 2))]
            print(f"  {lab}: take->done {ph[0]:.2f} | dep->first data/vector in smem {ph[1]:.2f} | products+reductions {ph[2]:.2f} | fence+signal {ph[3]:.2f} us (medians)")
            print(f"  {lab}: n={m.sum():6d}  wait(dep-take) med {np.median(w) / 1e3:7.2f} p90 {np.percentile(w, 90) / 1e3:7.2f} us | "
                  f"work(end-dep) med {np.median(p) / 1e3:6.2f} p90 {np.percentile(p, 90) / 1e3:6.2f} max {p.max() / 1e3:6.2f} us | queue depth mean {nq[m].mean():.2f}")
        # progress of the ticket frontier: when was ticket k finished
        order = np.argsort(end)
        idx = np.flatnonzero(ok)[order]
        for q in (0.1, 0.25, 0.5, 0.75, 0.9, 1.0):
            k = int(q * (len(order) - 1))
            print(f"  {int(q * 100):3d} % of the tickets done at {(end[order][k] - t0) / 1e3:8.1f} us (ticket index {idx[k]})")
        # the dependency chain: diagonal tickets in the order they became ready — gap between consecutive ones in the tail
        dd = np.sort(dep[isd])
        if len(dd) > 20:
            tail = np.diff(dd[-min(200, len(dd)):])
            print(f"  last {len(tail)} diagonal tickets: ready every {np.median(tail) / 1e3:.2f} us (median), {tail.sum() / 1e3:.1f} us in total")


if __name__ == "__main__":
    main()
