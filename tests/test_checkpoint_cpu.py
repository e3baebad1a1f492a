"""CPU: reference checkpoints load into our modules and ours load into the reference's."""
import os

import pytest
import torch

from oracle.ref_loader import load_reference_classes, reference_available


def test_roundtrip_with_module_prefix_and_head_drop(tmp_path):
    import maskunet_b200
    from maskunet_b200 import checkpoint as ck
    torch.manual_seed(0)
    a = maskunet_b200.UNet(3, 19)
    path = os.path.join(tmp_path, "dp.pth")
    ck.export_reference_checkpoint(a, path, data_parallel_prefix=True)        # what a DataParallel reference run saves
    assert all(k.startswith("module.") for k in torch.load(path))
    b = maskunet_b200.UNet(3, 19)
    missing, unexpected = ck.load_reference_checkpoint(b, path)
    assert not missing and not unexpected
    for (k, v), (_, w) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(v, w), k
    # cross-task transfer (city_semantic.py:335-338): another class count, head dropped, strict=False
    c = maskunet_b200.UNet(3, 150)
    missing, unexpected = ck.load_reference_checkpoint(c, path, drop_prefixes=("final_layer.",), strict=False)
    assert missing and all(k.startswith("final_layer.") for k in missing) and not unexpected
    assert torch.equal(c.state_dict()["bottom2.conv_block.0.weight"], a.state_dict()["bottom2.conv_block.0.weight"])


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("script,kw", [("ade_semantic", {}), ("city_instance", {"embed_dim": 16})])
def test_reference_state_dict_loads_both_ways(script, kw, tmp_path):
    import maskunet_b200
    from maskunet_b200 import checkpoint as ck
    ref = load_reference_classes(script)
    torch.manual_seed(1)
    r = ref.UNet(3, 19, **kw) if kw else ref.UNet(3, 19)
    path = os.path.join(tmp_path, "ref.pth")
    torch.save({"module." + k: v for k, v in r.state_dict().items()}, path)   # ade_semantic.py:344 under DataParallel
    ours = maskunet_b200.InstanceUNet(3, 19, **kw) if kw else maskunet_b200.UNet(3, 19)
    missing, unexpected = ck.load_reference_checkpoint(ours, path)
    assert not missing and not unexpected
    assert list(ours.state_dict().keys()) == list(r.state_dict().keys())
    out = os.path.join(tmp_path, "ours.pth")
    ck.export_reference_checkpoint(ours, out)
    r2 = ref.UNet(3, 19, **kw) if kw else ref.UNet(3, 19)
    r2.load_state_dict(torch.load(out))                                        # the reference's own loader accepts it
    for (k, v), (_, w) in zip(r.state_dict().items(), r2.state_dict().items()):
        assert torch.equal(v, w), k
