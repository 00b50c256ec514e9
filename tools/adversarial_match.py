"""K3 on adversarial seasons (BASELINE configs[1] shapes: 28 episodes, 2,413 / 1,201 hashes, 378 pairs,
openings + endings): does the adaptive match kernel (the default) ever lose to the exhaustive one?

The adaptive kernel tests 4 rows of every 32-row word per stage and leaves when no diagonal
survives; unrelated uniform hashes die in the first stage.  These seasons are built to keep
diagonals alive:
  random        uniform hashes + one planted 90 s run per list (the bench workload)
  correlated    every hash = the previous one with 3 bit flips (stationary background)
  silence60     60 s of digital silence (one constant hash, Chromaprint's 627964279) at a random
                place in every list: a 244 x 244 block of matching cells in every table
  jingle20      a 10 s jingle (40 hashes) repeated 20 times in every list: thousands of 40-cell
                runs, all below the 20 s minimum
  quiet_half    the second half of every list is silence
Each season is also matched by the exhaustive kernel (NB200_OPT_MATCH_DENSE) and the general
kernel (NB200_OPT_FORCE_GENERAL_MATCH); the three run lists must be identical.  With --oracle the
first 6 videos of each season are checked against the CPU oracle as well.
Prints one JSON object."""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from needle_b200 import engine, synth  # noqa: E402
from needle_b200._lib import OPT_FORCE_GENERAL_MATCH, OPT_MATCH_DENSE  # noqa: E402

N, NO, NE = 28, 2413, 1201


def make(kind, seed=7, n=N, no=NO, ne=NE):
    return synth.make_adversarial_season(kind, n, no, ne, seed=seed)


def main():
    ctx = engine.Context(0)
    p = engine.match_params(include_endings=True)
    out = {}
    kinds = ["random", "correlated", "silence60", "jingle20", "quiet_half"]
    for kind in kinds:
        s = make(kind)
        hs = engine.HashSet.upload(ctx, s.hashes, s.ts_ns, s.seg_offset)
        res = {}
        ref = None
        for name, opt in (("adaptive", None), ("dense", OPT_MATCH_DENSE), ("general", OPT_FORCE_GENERAL_MATCH)):
            if opt is not None:
                ctx.set_option(opt, 1)
            best, runs = None, None
            for it in range(4):
                rs = hs.match(p)
                ms = ctx.last_kernel_ms()["match"]
                best = ms if best is None else min(best, ms)
                if it == 0:
                    runs = rs.download()
                    n_runs, n_cells = rs.count()
                rs.free()
            if opt is not None:
                ctx.set_option(opt, 0)
            if ref is None:
                ref = runs
            res[name + "_ms"] = best
            res[name + "_equal"] = bool(np.array_equal(runs, ref))
        res["runs"] = int(n_runs)
        res["cells"] = int(n_cells)
        res["adaptive_over_dense"] = res["adaptive_ms"] / res["dense_ms"]
        hs.free()
        if "--oracle" in sys.argv:
            from oracle import oracle as orc
            from tests import helpers as H
            small = make(kind, n=6, no=700, ne=400)
            got = ctx.match_pairs(small.hashes, small.ts_ns, small.seg_offset, p)
            want = H.oracle_pair_runs(orc, small, include_endings=True)
            res["oracle_equal_small"] = H.runs_as_rows(got) == want
        out[kind] = res
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
