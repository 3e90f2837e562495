import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
from simuscop_b200 import cuda_binding, planfile, oracle_binding
wd = "/tmp/dbg_stat"; os.makedirs(wd, exist_ok=True)
helpers.SCENARIOS["stat"] = dict(lengths=[400000], names=["chr1"], profile="XTen", layout="PE", coverage=int(sys.argv[1]) if len(sys.argv) > 1 else 30, insertSize=300)
scn = helpers.build_scenario("stat", wd)
plans, out = helpers.run_reference_philox(scn)
plan = planfile.read_plan(plans[0])
r1p, r2p = helpers.sample_files(out, plan, 0, scn)
r1, r2 = helpers.read_file(r1p), helpers.read_file(r2p)
g = cuda_binding.Generator(0)
g.load_plan(plan, scn["seed"])
f1, f2 = g.generate()
d1, d2 = helpers.first_diff(f1, r1), helpers.first_diff(f2, r2)
print("pairs", g.emitted, "len", len(f1), len(r1), "first diff", d1, d2)
for f, r, d in ((f1, r1, d1), (f2, r2, d2)):
    if d >= 0:
        s = r.rfind(b"\n@", 0, d) + 1
        print("REF :", r[s:s + 700].decode(errors="replace"))
        s2 = f.rfind(b"\n@", 0, d) + 1
        print("OURS:", f[s2:s2 + 700].decode(errors="replace"))
        # count differing records
        ra, fa = r.split(b"\n"), f.split(b"\n")
        nd = sum(1 for x, y in zip(ra, fa) if x != y)
        print("differing lines", nd, "of", len(ra))
