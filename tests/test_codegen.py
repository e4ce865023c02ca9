"""CPU check of what the compiler made of the field products (tools/codegen_check.py; no GPU needed, reads the SASS of the built objects).
ptxas allocates registers over the whole call graph of a kernel: one caller that holds many field elements by value makes EVERY field
product of that kernel ~30 % longer (moves and spills) without any test failing -- it cost the pairing kernel 10 % and the G2 MSM a third
of its speed before it was noticed.  This test keeps the watched kernels within budget.  Skipped when the library has
not been built in this tree."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "kyber_b200", "csrc")


@pytest.mark.parametrize("obj", ["b2k_pairing.o", "b2k_g2.o", "b2k_api.o"])
def test_field_products_of_the_watched_kernels_are_compiled_clean(obj):
    path = os.path.join(CSRC, obj)
    if not os.path.exists(path) or shutil.which("cuobjdump") is None:
        pytest.skip("object not built here / cuobjdump not available")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "codegen_check.py"), path], capture_output=True, text=True)
    assert r.returncode == 0 and "over budget" not in r.stdout, r.stdout[-2000:]
    assert "276 wide" in r.stdout, "no 12-limb product found: the parser no longer understands the SASS listing"
