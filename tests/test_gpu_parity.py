"""GPU parity: the CUDA path (through the C ABI) against the Philox-instrumented reference
and the C oracle, byte for byte."""
import pytest

import helpers
from simuscop_b200 import planfile

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gen(built):
    from simuscop_b200 import cuda_binding
    g = cuda_binding.Generator(0)
    yield g
    g.close()


@pytest.mark.parametrize("name", ["pe_xten", "se_gaiix", "pe_tiny", "pe_variants", "pe_wes", "se_tumor"])
def test_fastq_bit_exact_vs_instrumented_reference(name, gen, workdir):
    scn = helpers.build_scenario(name, workdir)
    plans, out = helpers.run_reference_philox(scn)
    assert plans
    for i, pf in enumerate(plans):
        plan = planfile.read_plan(pf)
        r1p, r2p = helpers.sample_files(out, plan, i, scn)
        r1, r2 = helpers.read_file(r1p), helpers.read_file(r2p)
        gen.load_plan(plan, scn["seed"])
        f1, f2 = gen.generate()
        assert helpers.first_diff(f1, r1) == -1, "file 1 differs at byte %d" % helpers.first_diff(f1, r1)
        assert helpers.first_diff(f2, r2) == -1, "file 2 differs at byte %d" % helpers.first_diff(f2, r2)
        assert len(r1) > 0
