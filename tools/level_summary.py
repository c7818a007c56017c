"""Dev tool (CPU only): summarise a PB200_PROFILE_VERBOSE log (per-level, per-kernel-kind CUDA-event times of one
serialised factorization, written by tools/run_case.py under PB200_PROFILE=1 PB200_PROFILE_VERBOSE=1) — how much of
the factorization is the fixed launch chain of thin levels.  usage: python tools/level_summary.py LOG [--thin=TILES]"""
import re
import sys
from collections import defaultdict

KIND = {0: "diag", 1: "trsm", 2: "ext-update", 3: "in-panel update", 5: "fan-in", 6: "diag completion (LU)"}
PAT = re.compile(r"lvl\s+(\d+) kind (\d+) tasks\s+(\d+) tiles\s+(\d+) nbmax\s+(\d+) :\s+([0-9.]+) ms")


def main():
    path = sys.argv[1]
    thin = next((int(a.split("=")[1]) for a in sys.argv if a.startswith("--thin=")), 2)
    runs, cur, last = [], None, -1
    for line in open(path):
        m = PAT.search(line)
        if not m:
            continue
        lvl, kind, tasks, tiles, nb, ms = int(m[1]), int(m[2]), int(m[3]), int(m[4]), int(m[5]), float(m[6])
        if lvl < last or cur is None:
            cur = defaultdict(list); runs.append(cur)
        last = lvl
        cur[lvl].append((kind, tasks, tiles, ms))
    if not runs:
        raise SystemExit("no level lines found")
    lv = runs[-1]                                   # the last factorization of the log (warm)
    total = sum(ms for rows in lv.values() for *_, ms in rows)
    launches = sum(len(rows) for rows in lv.values())
    # a thin level: its panel work (diag) covers at most `thin` cblks — the chains of the top separators
    thin_lv = [l for l, rows in lv.items() if max(t for k, t, *_ in rows if k == 0) <= thin]
    per_kind = defaultdict(lambda: [0.0, 0])
    for l in thin_lv:
        for k, _, _, ms in lv[l]:
            per_kind[k][0] += ms; per_kind[k][1] += 1
    t_thin = sum(v[0] for v in per_kind.values())
    print(f"{path}: {len(lv)} levels, {launches} launches, {total:.2f} ms serialised")
    print(f"thin levels (<= {thin} cblks): {len(thin_lv)} levels, {sum(len(lv[l]) for l in thin_lv)} launches, "
          f"{t_thin:.2f} ms = {100 * t_thin / total:.1f} % of the serialised time, {1e3 * t_thin / max(len(thin_lv), 1):.0f} us per level")
    for k in sorted(per_kind):
        ms, n = per_kind[k]
        print(f"  {KIND.get(k, k):22s} {n:5d} launches  {ms:7.3f} ms  {1e3 * ms / n:6.1f} us each")


if __name__ == "__main__":
    main()
