"""Reference checkpoint format round trip (run_train.py:369-380 <-> tools/create_model_condition.py:62-89), CPU."""
import torch

from mofanerf_b200 import B200Renderer, nets
from mofanerf_b200.checkpoint import TRAIN_KEYS, load_train_checkpoint, save_train_checkpoint
from oracle import mofa_oracle as O
from oracle import ref_loader


def test_round_trip_and_reference_compatibility(tmp_path):
    c, f, _ = nets.build_nets(3, 256, 8, 256, 10)
    r = B200Renderer(expCodesLen=30)
    p = str(tmp_path / "000100.tar")
    save_train_checkpoint(p, 100, torch.nn.DataParallel(c), f, r)
    ck = torch.load(p, map_location="cpu", weights_only=False)
    assert set(TRAIN_KEYS) <= set(ck)
    r2 = B200Renderer(expCodesLen=30)
    c2, f2, step = load_train_checkpoint(p, r2)
    assert step == 100 and (c2.W, c2.D, f2.W, f2.D) == (256, 8, 256, 10)
    for a, b in ((c, c2), (f, f2), (r.texEncoder, r2.texEncoder), (r.idSpecificMod, r2.idSpecificMod)):
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa) == list(sb) and all(torch.equal(sa[k], sb[k]) for k in sa)
    assert all(torch.equal(x, y) for x, y in zip(r.expCodes_Sigma, r2.expCodes_Sigma))
    # the oracle's (== the reference's) NeRF accepts the saved state_dicts unchanged
    oc = O.NeRF(8, 256, 93, 27, 256, 50)
    oc.load_state_dict(ck["network_fn_state_dict"])
    if ref_loader.available():
        ref = ref_loader.load()
        rn = ref.model.NeRF(D=10, W=256, input_ch_shapeCodes=50, input_ch_textureCodes=256, input_ch=93, output_ch=5,
                            skips=[4], input_ch_views=27, use_viewdirs=True)
        rn.load_state_dict(ck["network_fine_state_dict"])
