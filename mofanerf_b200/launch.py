"""Run an UNCHANGED reference entry script (run_fit.py, run_train.py, render_refine_trainSet.py) on the B200 engine:

    python -m mofanerf_b200.launch [--shim-missing] [--shard] /path/to/mofanerf/run_fit.py --filePath ... --renderType rendering
    torchrun --nproc-per-node 8 -m mofanerf_b200.launch --shard /path/to/mofanerf/run_fit.py ...

What it does before handing control to the script (runpy, `__name__ == "__main__"`):

1. Initialises CUDA on this process's GPU (LOCAL_RANK, default 0) FIRST.  Every reference script starts with
   `os.environ['CUDA_VISIBLE_DEVICES'] = '0'` (run_fit.py:3, run_train.py:12, render_refine_trainSet.py:2); CUDA reads
   that variable once, when the runtime initialises, so an assignment made after initialisation is inert: rank r keeps
   GPU r as its current device, and `torch.set_default_tensor_type('torch.cuda.FloatTensor')` (run_fit.py:438) makes
   that device the default.  The script itself is not modified.
2. Puts the reference tree (the script's directory unless --reference-root is given) on sys.path, changes into it (the
   scripts read ./configs/... relative to the working directory) and installs the engine:
   models.render_class.myRenderer = mofanerf_b200.B200Renderer (mofanerf_b200.install()).
3. --shim-missing: provides minimal stand-ins for third-party modules the scripts import at module level but that are
   not installed (imageio -> PIL, configargparse -> argparse + "key = value" config files, matplotlib.pyplot and dlib ->
   empty modules, removed numpy aliases np.int / np.long / np.float / np.bool).  Installed modules are never replaced.
4. Under torchrun (WORLD_SIZE > 1): torch.distributed over NCCL; with --shard every render()/render_fitting() call
   renders this rank's contiguous ray range and one all-gather rebuilds the maps (MOFA_B200_SHARD=1).
"""
from __future__ import annotations

import argparse
import importlib
import os
import runpy
import sys
import types


def neutralise_device_pin(local_rank: int) -> bool:
    """Initialise CUDA on `local_rank` now, so that a later `os.environ['CUDA_VISIBLE_DEVICES'] = '0'` has no effect."""
    import torch
    if not torch.cuda.is_available():
        return False
    torch.cuda.set_device(local_rank)
    torch.cuda.init()
    torch.zeros(1, device=f"cuda:{local_rank}")      # forces context creation on the device
    torch.cuda.device_count()                         # cached by torch once CUDA is initialised: later calls do not re-read the env
    return True


def _have(name: str, attr: str = None) -> bool:
    """Importable — and, where `attr` is given, actually providing it (an empty placeholder module does not count)."""
    try:
        mod = importlib.import_module(name)
    except Exception:
        return False
    return attr is None or hasattr(mod, attr)


def _imageio_shim() -> types.ModuleType:
    import numpy as np
    from PIL import Image
    m = types.ModuleType("imageio")

    def imread(path, *a, **k):
        return np.asarray(Image.open(path))

    def imwrite(path, img, *a, **k):
        Image.fromarray(np.asarray(img)).save(path)

    m.imread, m.imwrite, m.imsave = imread, imwrite, imwrite
    m.__mofa_shim__ = True
    return m


def _configargparse_shim() -> types.ModuleType:
    m = types.ModuleType("configargparse")

    class ArgumentParser(argparse.ArgumentParser):
        """argparse + the one configargparse feature tools/config_parser.py uses: an `is_config_file=True` option whose
        file holds `key = value` lines (comments after '#'); command-line options override the file."""

        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self._config_opts = []
            self._store_true = set()

        def add_argument(self, *names, **kw):
            if kw.pop("is_config_file", False):
                self._config_opts.append((names, kw.get("default")))
            act = super().add_argument(*names, **kw)
            if isinstance(act, argparse._StoreTrueAction):
                self._store_true.add(act.dest)
            return act

        def _file_args(self, argv):
            path = None
            for names, default in self._config_opts:
                path = default
                for i, a in enumerate(argv):
                    if a in names and i + 1 < len(argv):
                        path = argv[i + 1]
                    for n in names:
                        if a.startswith(n + "="):
                            path = a.split("=", 1)[1]
            out = []
            if path and os.path.exists(path):
                for line in open(path):
                    line = line.split("#", 1)[0].strip()
                    if "=" not in line:
                        continue
                    key, val = [x.strip() for x in line.split("=", 1)]
                    if key in self._store_true:
                        if val.lower() in ("true", "1", "yes"):
                            out.append("--" + key)
                    else:
                        out += ["--" + key, val]
            return out

        def parse_known_args(self, args=None, namespace=None):
            argv = list(sys.argv[1:] if args is None else args)
            return super().parse_known_args(self._file_args(argv) + argv, namespace)

    m.ArgumentParser = ArgumentParser
    m.ArgParser = ArgumentParser
    m.__mofa_shim__ = True
    return m


def shim_missing_modules() -> list:
    """Stand-ins for modules the reference scripts import at module level but that are absent here.  Returns the names
    that were shimmed."""
    done = []
    makers = {"imageio": (_imageio_shim, "imwrite"), "configargparse": (_configargparse_shim, "ArgumentParser")}
    for name, (make, attr) in makers.items():
        if not _have(name, attr):
            shim = make()
            if name in sys.modules:          # a placeholder other code may already hold by reference: fill it in place
                for k, v in shim.__dict__.items():
                    if not k.startswith("__") or k == "__mofa_shim__":
                        setattr(sys.modules[name], k, v)
            else:
                sys.modules[name] = shim
            done.append(name)
    for name in ("dlib", "matplotlib"):
        if not _have(name):
            mod = types.ModuleType(name)
            mod.__mofa_shim__ = True
            sys.modules[name] = mod
            done.append(name)
            if name == "matplotlib":
                plt = types.ModuleType("matplotlib.pyplot")
                plt.__mofa_shim__ = True
                mod.pyplot = plt
                sys.modules["matplotlib.pyplot"] = plt
    import numpy as np
    for alias, typ in (("int", int), ("long", int), ("float", float), ("bool", bool)):
        if not hasattr(np, alias):          # removed in numpy 1.24 / 2.0; the scripts were written for 1.19
            setattr(np, alias, typ)
            done.append("numpy." + alias)
    return done


def main(argv=None) -> None:
    ap = argparse.ArgumentParser(prog="python -m mofanerf_b200.launch", description=__doc__.split("\n\n")[0])
    ap.add_argument("--reference-root", default=None, help="reference tree (default: the script's directory)")
    ap.add_argument("--shim-missing", action="store_true", help="stand-ins for missing imageio / configargparse / ...")
    ap.add_argument("--shard", action="store_true", help="ray-shard every image across the ranks (MOFA_B200_SHARD=1)")
    ap.add_argument("--async-png", action="store_true",
                    help="render_path returns when the frames are in host memory; PNG files are written in the background "
                         "(MOFA_B200_ASYNC_PNG=1) and joined at exit")
    ap.add_argument("--no-chdir", action="store_true", help="do not change into the reference tree")
    ap.add_argument("script")
    ap.add_argument("script_args", nargs=argparse.REMAINDER)
    a = ap.parse_args(argv)

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    have_gpu = neutralise_device_pin(local_rank)

    script = os.path.abspath(a.script)
    root = os.path.abspath(a.reference_root) if a.reference_root else os.path.dirname(script)
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (here, root):
        if p not in sys.path:
            sys.path.insert(0, p)
    if a.shim_missing:
        shimmed = shim_missing_modules()
        if shimmed:
            print(f"[mofanerf_b200.launch] stand-ins for: {', '.join(shimmed)}")
    if not a.no_chdir:
        os.chdir(root)

    import mofanerf_b200
    mofanerf_b200.install()

    if world > 1:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("nccl" if have_gpu else "gloo",
                                    **({"device_id": torch.device("cuda", local_rank)} if have_gpu else {}))
    if a.shard:
        os.environ["MOFA_B200_SHARD"] = "1"
    if a.async_png:
        os.environ["MOFA_B200_ASYNC_PNG"] = "1"

    sys.argv = [script] + list(a.script_args)
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
