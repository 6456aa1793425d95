"""GPU tier: the device-side input pipeline (pab_prepare_clouds, CloudFeeder, make_descs) against the reference's own
normalisation (golden vectors) and the oracle restatement of SceneDataSet.get_pc."""
import os

import numpy as np
import pytest
import torch

import util
from oracle import prepare
from patchaugnet_b200 import input_pipeline as ip

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_prepare_batch_matches_the_reference_normalisation():
    g = np.load(os.path.join(util.GOLDEN, "prepare_ref.npz"))
    for i in range(3):
        raw = torch.from_numpy(g[f"raw{i}"])[None].to(DEV)
        for zoom in (True, False):
            out, metas = ip.prepare_batch(raw, g[f"offset{i}"], normalize=True, zoom=zoom, return_norm_meta=True)
            want = g[f"pc{i}_zoom{int(zoom)}"]
            assert out.shape == (1, 1, want.shape[0], 3) and out.dtype == torch.float32
            # float64 arithmetic on both sides, cast last: equal up to one float32 ulp where the float64 sums round differently
            got = out[0, 0].cpu().numpy()
            assert np.abs(got - want.astype(np.float32)).max() <= 2e-7 * max(1.0, np.abs(want).max())
            assert (got == want.astype(np.float32)).mean() > 0.999
            assert abs(metas[0]["scale"] - float(g[f"scale{i}_zoom{int(zoom)}"])) <= 1e-12 * max(1.0, float(g[f"scale{i}_zoom{int(zoom)}"]))
            assert np.abs(metas[0]["trans"] - g[f"trans{i}_zoom{int(zoom)}"].reshape(-1)).max() < 1e-9
        out = ip.prepare_batch(raw, g[f"offset{i}"], normalize=False)
        assert np.array_equal(out[0, 0].cpu().numpy(), (g[f"raw{i}"] - g[f"offset{i}"]).astype(np.float32))


def test_prepare_batch_sizes_and_errors():
    rng = np.random.default_rng(3)
    for n in (2, 33, 4096, 8192):
        raw = rng.normal(size=(5, n, 3)) * 10 + 100
        out = ip.prepare_batch(torch.from_numpy(raw).to(DEV), [100.0, 100.0, 100.0], normalize=True, zoom=True)
        for b in range(5):
            want, _ = prepare.get_pc(raw[b], np.array([100.0, 100.0, 100.0]), normalize=True, zoom=True)
            assert np.abs(out[b, 0].cpu().numpy() - want.astype(np.float32)).max() <= 2e-7
        assert abs(float(out[0, 0].norm(dim=1).max()) - 1.0) < 1e-6                  # unit ball after zoom
    with pytest.raises(ValueError):
        ip.prepare_batch(torch.zeros(2, 8193, 3, dtype=torch.float64, device=DEV))
    with pytest.raises(Exception):
        ip.prepare_batch(torch.zeros(2, 16, 3, dtype=torch.float64))                 # CPU tensor: no fallback


def test_make_descs_from_bin_files_matches_the_host_prepared_forward(tmp_path):
    """.bin files -> CloudFeeder (pinned staging, copy stream, device normalisation) -> fused engine, ragged last batch, vs
    the reference order of operations: load, offset, normalise on the host, upload, forward."""
    cfg = dict(util.PATCHAUGNET_CFG, SAMPLING=[256, 64, 16], MAX_SAMPLES=[64, 256, 1024])
    net = util.build_network(DEV, cfg=cfg)
    rng = np.random.default_rng(9)
    offset = np.array([5.0e4, 7.0e5, 10.0])
    files, host = [], []
    for i in range(11):
        raw = rng.normal(size=(1024, 3)) * np.array([30.0, 30.0, 4.0]) + offset + rng.normal(size=3) * 50
        f = tmp_path / f"{i}.bin"
        raw.astype(np.float64).tofile(f)
        files.append(str(f))
        pc, _ = prepare.get_pc(prepare.load_pc_file(str(f)), offset, normalize=True, zoom=True)
        host.append(torch.from_numpy(pc).float())
    with torch.no_grad():
        want = torch.cat([net(torch.stack(host[i:i + 4]).unsqueeze(1).to(DEV), return_feat=False) for i in range(0, 11, 4)])
        got = ip.make_descs(net, files, batch_size=4, num_points=1024, global_offset=offset, normalize=True, zoom=True, super_chunk=2)
        torch.cuda.synchronize()
    assert got.shape == (11, 256)
    assert (got - want).abs().max().item() < 1e-4
    # in-memory float32 sources, no normalisation: the feeder is a plain cast + batcher
    clouds = [util.synthetic_cloud(i, 1024).numpy() for i in range(8)]
    with torch.no_grad():
        got = ip.make_descs(net, clouds, batch_size=4, num_points=1024, dtype=np.float32)
        want = net(torch.from_numpy(np.stack(clouds)).unsqueeze(1).to(DEV), return_feat=False)
    assert torch.equal(got, want)
