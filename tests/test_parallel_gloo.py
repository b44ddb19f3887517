"""CPU test of the N>1 path (world_size 2, gloo): weight broadcast from rank 0, batch
sharding, per-rank compute (the oracle stands in for the device kernels here), and
host-side concatenation must reproduce the single-process result exactly."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import lce_testlib as L
from compute_engine_b200 import parallel as P


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_layers():
    rng = np.random.default_rng(3)
    layers = []
    for c in (64, 96):
        layers.append({"filter": rng.integers(-2**31, 2**31, (c, 3, 3, L.cdiv(c, 32)),
                                               dtype=np.int64).astype(np.int32),
                       "mul": rng.uniform(0.01, 1.5, c).astype(np.float32),
                       "bias": rng.uniform(-1, 1, c).astype(np.float32)})
    layers.append({"filter": rng.integers(-2**31, 2**31, (32, 1, 1, 3), dtype=np.int64).astype(np.int32),
                   "thresholds": rng.integers(30, 60, 32).astype(np.int32)})
    return layers


def _run_model(layers, x):
    """x: float [b,6,6,64] -> conv(64->64) -> conv? (separate 96-ch input) ... keep it simple:
    layer0 on x, bitpacked layer2 on a 96-channel tensor derived from x."""
    b = x.shape[0]
    d0 = L.BconvDesc(b, 6, 6, 64, 3, 3, 64, 1, 1, 1, 1, 1, L.PADDING_SAME, 1, L.ACT_RELU,
                     L.OUT_FLOAT, 1.0, 0)
    y = L.bconv2d(d0, L.quantize(x), layers[0]["filter"], layers[0]["mul"], layers[0]["bias"])
    x96 = np.concatenate([y, x[..., :32]], axis=-1)
    d2 = L.BconvDesc(b, 6, 6, 96, 1, 1, 32, 1, 1, 1, 1, 1, L.PADDING_VALID, 1, L.ACT_NONE,
                     L.OUT_BITPACKED, 1.0, 0)
    return L.bconv2d(d2, L.quantize(x96), layers[2]["filter"], thr=layers[2]["thresholds"])


def _worker(rank, world, port, batch, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    layers = _make_layers() if rank == 0 else None
    got, blob = P.broadcast_model(layers, torch.device("cpu"))          # the ONE collective
    np_layers = [{k: v.numpy() for k, v in lay.items()} for lay in got]
    x = np.random.default_rng(11).standard_normal((batch, 6, 6, 64)).astype(np.float32)
    lo, hi = P.shard_range(rank, world, batch)
    local = _run_model(np_layers, x[lo:hi])                              # no step collective
    full = P.gather_outputs(torch.from_numpy(local), batch)
    if rank == 0:
        q.put((full.numpy(), int(blob.numel())))
    dist.destroy_process_group()


def test_shard_range_partitions_the_batch():
    for world in (1, 2, 4, 8):
        for batch in (0, 1, 7, 256, 1024, 1025):
            spans = [P.shard_range(r, world, batch) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert P.shard_range(3, 8, 1024) == (384, 512)      # config 3: 128 images per GPU


def test_weight_blob_roundtrip():
    layers = _make_layers()
    blob, manifest = P.pack_weight_blob(layers)
    back = P.unpack_weight_blob(torch.from_numpy(blob), manifest)
    for a, b in zip(layers, back):
        assert a.keys() == b.keys()
        for k in a:
            assert np.array_equal(a[k], b[k].numpy()) and a[k].dtype == b[k].numpy().dtype


def test_world_size_2_matches_single_process():
    batch = 5                                  # ragged: 3 + 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    full, blob_words = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x = np.random.default_rng(11).standard_normal((batch, 6, 6, 64)).astype(np.float32)
    want = _run_model(_make_layers(), x)
    assert np.array_equal(full, want)
    assert blob_words == sum(int(np.prod(v.shape)) for lay in _make_layers() for v in lay.values())
