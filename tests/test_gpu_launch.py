"""`python -m mofanerf_b200.launch` on the GPU box: a script that pins CUDA_VISIBLE_DEVICES='0' on its first line — as
run_fit.py:3, run_train.py:12 and render_refine_trainSet.py:2 do — still runs on the GPU its LOCAL_RANK names, with the
engine installed as models.render_class.myRenderer.  (The reference tree does not exist on this box: the script and the
two-line `models` package are written to a temporary directory; the unchanged run_fit.py itself is exercised by
tests/test_scripts_cpu.py in the build container.)"""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = '''import os
os.environ['CUDA_VISIBLE_DEVICES'] = '0'
import torch
import models.render_class as rc
r = rc.myRenderer(expCodesLen=30)
torch.set_default_tensor_type('torch.cuda.FloatTensor')
x = torch.zeros(3)
print("RESULT", type(r).__name__, torch.cuda.device_count(), torch.cuda.current_device(), x.device.index)
'''


def test_launcher_neutralises_the_device_pin(tmp_path):
    (tmp_path / "models").mkdir()
    (tmp_path / "models" / "__init__.py").write_text("")
    (tmp_path / "models" / "render_class.py").write_text("class myRenderer:\n    pass\n")
    (tmp_path / "script.py").write_text(SCRIPT)
    n_gpu = torch.cuda.device_count()
    rank = 1 if n_gpu >= 2 else 0
    env = dict(os.environ, LOCAL_RANK=str(rank), PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    env.pop("CUDA_VISIBLE_DEVICES", None)
    out = subprocess.run([sys.executable, "-m", "mofanerf_b200.launch", str(tmp_path / "script.py")], env=env, cwd=ROOT,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("RESULT")][-1].split()
    assert line[1] == "B200Renderer"
    assert int(line[2]) == n_gpu, "the pin written by the script must not have hidden the other GPUs"
    assert int(line[3]) == rank and int(line[4]) == rank
