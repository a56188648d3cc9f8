"""Collects the per-case parity numbers the GPU tests measure and writes them to a JSON file at session end
(profiles/parity_r02.json by default, $MOFA_PARITY_JSON to redirect) — `pytest -q` hides the [parity] prints, the
committed table does not.  Test infrastructure only."""
import json
import os
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_records = {}


def record(case: str, **metrics) -> None:
    clean = {}
    for k, v in metrics.items():
        try:
            clean[k] = None if v is None else round(float(v), 9)
        except (TypeError, ValueError):
            clean[k] = str(v)
    _records.setdefault(case, {}).update(clean)


def flush() -> None:
    if not _records:
        return
    path = os.environ.get("MOFA_PARITY_JSON", os.path.join(ROOT, "profiles", "parity_r02.json"))
    doc = {"what": "engine (CUDA, C ABI) vs reference fixtures / oracle: per-case max / mean |delta| and PSNR, written by "
                   "`pytest -m gpu` (tests/parity_log.py)",
           "tolerance": "SURVEY.md §7 / BASELINE.md §4.4: max|d rgb| <= 3e-2, mean <= 3e-3, PSNR >= 45 dB",
           "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "cases": _records}
    try:
        import torch
        if torch.cuda.is_available():
            doc["gpu"] = torch.cuda.get_device_name(0)
    except Exception:
        pass
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            json.dump(doc, f, indent=1, sort_keys=True)
    except OSError:
        pass
