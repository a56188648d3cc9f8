"""bench.py host-side contract (no GPU): the reference arm prints one JSON line with the keys the driver reads, the
pipeline labels follow SURVEY.md §8(d) (FULL vs COARSE64), and the FLOP accounting matches the oracle's enumeration of
the reference's nn.Linear layers."""
import json
import os
import subprocess
import sys
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from oracle import mofa_oracle as O  # noqa: E402


def _args(**kw):
    d = dict(H=800, W=800, n_samples=64, n_importance=64, gpus=1)
    d.update(kw)
    return types.SimpleNamespace(**d)


def test_default_metric_is_the_baseline_metric():
    a = _args()
    assert bench.metric_name(a) == "rays/sec at 800x800x64 samples (FULL: 64 coarse + 128 fine evaluations/ray)"
    assert bench.fine_evals(a) == 128
    assert "FULL pipeline" in bench.workload_config(a)["workload"]


def test_coarse64_pipeline_has_no_fine_pass():
    a = _args(n_importance=0)
    assert bench.fine_evals(a) == 0
    assert "COARSE64" in bench.metric_name(a)
    assert "no fine pass" in bench.workload_config(a)["workload"]


def test_flop_constants_match_the_layer_enumeration():
    # coarse point 3 187 200, fine point 54 953 984, FULL ray 7 238.1 MFLOP
    assert bench.FLOP_COARSE_PT == 3187200.0
    assert bench.FLOP_FINE_PT == 54953984.0
    full = O.flops_per_ray()
    assert full == pytest.approx(64 * bench.FLOP_COARSE_PT + 128 * bench.FLOP_FINE_PT, rel=1e-12)
    assert full / 1e6 == pytest.approx(7238.1, abs=0.05)


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--H", "8", "--W", "8", "--n-samples", "8", "--n-importance", "8"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "rays/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1


def test_parity_stats_separates_opacity_gate_flips_from_arithmetic_error():
    """bench.parity_stats: a ray whose sigma at the last sample changes sign between the two renderings is an opacity-gate
    flip (the reference's relu(sigma) * 1e10 step, models/render_class.py:449) and is reported apart from the rest."""
    import torch
    g = torch.Generator().manual_seed(0)
    ref = torch.rand(200, 3, generator=g)
    got = ref + 1e-3 * torch.randn(200, 3, generator=g)
    acc_ref = torch.ones(200)
    acc = acc_ref.clone()
    sig_ref = torch.randn(200, generator=g).abs() + 0.5          # firmly positive everywhere ...
    sig = sig_ref.clone()
    sig_ref[7], sig[7] = -2e-4, 3e-4                            # ... except one ray that sits on the step
    got[7] += 0.6
    acc[7] = 0.9
    st = bench.parity_stats(got, acc, ref, acc_ref, sig, sig_ref)
    assert st["opacity_gate_flips"] == 1 and st["gate_flips_with_visible_effect"] == 1
    assert abs(st["gate_flip_max_abs_sigma_last_of_reference"] - 2e-4) < 1e-9
    assert st["max_abs_rgb"] > 0.5 and st["max_abs_rgb_excluding_gate_flips"] < 1e-2
    assert st["rays_over_3e-2"] == 1 and st["rays_over_3e-2_excluding_gate_flips"] == 0
    assert st["psnr_db_excluding_gate_flips"] > st["psnr_db"] + 10
    # without the sigmas a flip is a ray whose accumulated opacity differs by more than 0.5: this one (0.9 vs 1.0) is not
    st2 = bench.parity_stats(got, acc, ref, acc_ref)
    assert st2["opacity_gate_flips"] == 0 and "gate_flips_with_visible_effect" not in st2
