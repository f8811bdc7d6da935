"""Generate tests/golden/af_vectors.json by EXECUTING the reference's own arithmetic (AST-extracted from
/root/reference/src/telr/TELR_te.py, see tests/ref_arith.py).  Run in the build container only."""
import json
import os
import random
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests import ref_arith  # noqa: E402


def fmt(v):
    return None if v is None else (v if isinstance(v, int) else float(v))


def main():
    rnd = random.Random(20221101)
    cases = []
    shapes = [(8000, 3000, 7000), (8000, 300, 7000), (8000, 299, 7000), (8000, 3000, 3060), (8000, 3000, 3100), (8000, 3000, 3101),
              (8000, 3000, 7700), (8000, 3000, 7701), (900, 100, 800), (8000, 0, 5000), (8000, 3000, 8000), (5000, 350, 4650)]
    for (L, s, e) in shapes + [(rnd.randint(700, 12000),) * 1 + (0, 0) for _ in range(28)]:
        if e == 0:
            s = rnd.randint(0, L - 120)
            e = min(L, s + rnd.randint(20, 6000))
        for params in ((100, 200, 50, 50), (100, 200, 0, 50), (30, 10, 20, 5)):
            fl, fo, ti, to = params
            seed = rnd.randint(0, 1 << 30)
            dfw, drc = ref_arith.synth_depth(seed, L, s, e)
            fns = ref_arith.ref_functions({"fw": dfw, "rc": drc})
            try:
                fw = list(fns["get_te_cov"]("fw", "c", s, e, ti, to)) + list(fns["get_flank_cov"]("fw", "c", L, s, e, fl, fo))
                rc = list(fns["get_te_cov"]("rc", "c", L - e, L - s, ti, to)) + list(fns["get_flank_cov"]("rc", "c", L, L - e, L - s, fl, fo))
            except Exception as ex:       # statistics.StatisticsError on empty windows
                cases.append(dict(L=L, s=s, e=e, params=params, seed=seed, error=type(ex).__name__))
                continue
            # run the reference AF block on .freq files written the way the reference writes them
            row = ["chr", "100", "101"] + ["x"] * 11
            with tempfile.TemporaryDirectory() as td:
                f1, f2 = os.path.join(td, "a.freq"), os.path.join(td, "a.revcomp.freq")
                open(f1, "w").write("\t".join(row + [str(v) for v in fw]) + "\n")
                open(f2, "w").write("\t".join(row + [str(v) for v in rc]) + "\n")
                te_freq = ref_arith.ref_af_block(f1, f2)
            d = te_freq["chr_100_101"]
            cases.append(dict(L=L, s=s, e=e, params=params, seed=seed, fw_str=[str(v) for v in fw], rc_str=[str(v) for v in rc],
                              freq=d["freq"], freq_is_int=isinstance(d["freq"], int),
                              keys={k: d[k] for k in d if k != "freq"}))
    # the 7 known-answer vectors of SURVEY.md 8c through the reference AF block
    kav = []
    for fw, rc in [(("12", "13", "24", "25"), ("11", "12", "22", "23")), (("30", "30", "24", "24"), ("12", "12", "24", "24")),
                   (("37", "30", "24", "24"), ("12", "12", "24", "24")), (("26", "26", "24", "24"), ("27", "27", "24", "24")),
                   (("0", "0", "24", "24"), ("None", "3", "None", "4")), (("23.5", "1", "47", "1"), ("7", "1", "21.0", "1")),
                   (("5", "5", "16", "16"), ("5", "5", "None", "16"))]:
        row = ["chr", "100", "101"] + ["x"] * 11
        with tempfile.TemporaryDirectory() as td:
            f1, f2 = os.path.join(td, "a.freq"), os.path.join(td, "a.revcomp.freq")
            open(f1, "w").write("\t".join(row + list(fw)) + "\n")
            open(f2, "w").write("\t".join(row + list(rc)) + "\n")
            d = ref_arith.ref_af_block(f1, f2)["chr_100_101"]
        kav.append(dict(fw=fw, rc=rc, freq=d["freq"], freq_is_int=isinstance(d["freq"], int)))
    json.dump(dict(cases=cases, known_answers=kav, source="/root/reference/src/telr/TELR_te.py (AST-executed)"),
              open(os.path.join(HERE, "af_vectors.json"), "w"), indent=0)
    print(len(cases), "cases,", sum("error" in c for c in cases), "error cases,", len(kav), "known-answer vectors")


if __name__ == "__main__":
    main()
