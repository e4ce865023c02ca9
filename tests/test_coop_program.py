"""CPU check of the warp-cooperative pairing's program tables (tools/gen_coop_pairing.py --check): the tower / Miller / final-exponentiation
formulas over plain integers, and the ENCODED, scheduled, slot-allocated program interpreted numerically (reads of a round before its
writes, like the lanes of a warp), must both reproduce the oracle's pairing -- for the 1-pair program (Suite.Pair) and the 2-pair program
(ValidatePairing, incl. a true e(aG, bH) e(-abG, H) = 1 instance) of BLS12-381, bn254 and bn256; and the generated
kyber_b200/csrc/coop_program_*.inc must be the generator's current output."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_program_tables_reproduce_the_oracle_and_are_up_to_date():
    inc = os.path.join(ROOT, "kyber_b200", "csrc", "coop_program_bn256.inc")
    if not os.path.exists(inc):                              # generated at build time (git-ignored)
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_coop_pairing.py")], check=True, stdout=subprocess.DEVNULL)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_coop_pairing.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("encoded program ok") == 24 and "WRONG" not in r.stdout and "up to date" in r.stdout, r.stdout
