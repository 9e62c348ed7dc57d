"""CLI surface (reference fbs_mapper/map_circuit.py:92-188): flags, stats line, output files."""
import ast
import os
import subprocess
import sys

import pytest

from conftest import GOLD, ROOT, load_ref_mapped

REF_KEYS = ["nb_inp", "nb_linprod", "nb_bootstrap", "max_lut_size", "norm2_linprod", "nb_out", "filename", "type", "fbs_size",
            "mapper", "strict_fbs_size", "output", "output_lbf", "max_tt_size", "verbose", "max_fbs_size", "time"]


def run_cli(args):
    return subprocess.run([sys.executable, "-m", "tfhe_fbs_map_b200.map_circuit"] + args, cwd=ROOT, capture_output=True, text=True)


@pytest.mark.parametrize("circuit,p,mapper", [("full_adder", 15, "search"), ("aes_sbox", 15, "naive"), ("half_adder", 15, "basic"), ("ascon_lut", 17, "search")])
def test_cli_maps_and_writes_reference_outputs(tmp_path, circuit, p, mapper):
    e = next(x for x in load_ref_mapped() if x["circuit"] == circuit and x["p"] == p and x["mapper"] == mapper and not x.get("strict"))
    fbs, lbf = tmp_path / "o.fbs", tmp_path / "o.lbf"
    r = run_cli([os.path.join(GOLD, "blif", f"{circuit}.blif"), "--fbs_size", str(p), "--mapper", mapper, "--output", str(fbs),
                 "--output_lbf", str(lbf), "--exec", "none"])
    assert r.returncode == 0, r.stderr
    d = ast.literal_eval(r.stdout.strip().splitlines()[-1])         # what experiments/build_csv.py:24-25 parses
    assert list(d.keys()) == REF_KEYS
    for k, v in e["stats"].items():
        assert d[k] == v
    assert d["max_fbs_size"] == 2 * p and d["fbs_size"] == p and d["mapper"] == mapper
    assert lbf.read_text() == e["lbf"] and fbs.read_text() == e["fbs"]


def test_cli_strict_flag(tmp_path):
    r = run_cli([os.path.join(GOLD, "blif", "aes_sbox.blif"), "--fbs_size", "15", "--strict_fbs_size", "--exec", "none"])
    e = next(x for x in load_ref_mapped() if x["circuit"] == "aes_sbox" and x.get("strict"))
    d = ast.literal_eval(r.stdout.strip().splitlines()[-1])
    assert d["max_fbs_size"] == 15 and d["strict_fbs_size"] is True
    assert {k: d[k] for k in e["stats"]} == e["stats"]


def test_cli_without_gpu_maps_anyway_but_encrypted_exec_fails_loudly():
    """The reference-compatible invocation `map_circuit FILE --fbs_size N` has no GPU dependency in the reference
    (experiments/build_csv.py parses its last stdout line): without a usable GPU the cleartext self-check is skipped with a
    warning and the stats line is still printed.  `--exec b200` (the encrypted executor) has no CPU fallback and fails."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = run_cli([os.path.join(GOLD, "blif", "half_adder.blif"), "--fbs_size", "15"])
    assert r.returncode == 0 and "warning: GPU self-check unavailable" in r.stderr, r.stderr
    d = ast.literal_eval(r.stdout.strip().splitlines()[-1])
    assert d["nb_bootstrap"] == 2 and d["fbs_size"] == 15
    r = run_cli([os.path.join(GOLD, "blif", "half_adder.blif"), "--fbs_size", "15", "--exec", "b200"])
    assert r.returncode != 0 and "fbs error" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["clear", "b200"])
def test_cli_self_check_on_gpu(tmp_path, mode):
    r = run_cli([os.path.join(GOLD, "blif", "aes_sbox.blif"), "--fbs_size", "11", "--exec", mode, "--batch", "64" if mode == "b200" else "1000"])
    assert r.returncode == 0, r.stderr + r.stdout
    d = ast.literal_eval(r.stdout.strip().splitlines()[-1])
    assert d["nb_bootstrap"] == 38 and d["norm2_linprod"] == 35
