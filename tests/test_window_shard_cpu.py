"""Frame-sharded window (N > 1 ranks) — host-side logic on CPU with the gloo backend, world_size 2.

Checks the partition, the all-gather layout the attention kernel consumes (`kv_chunks` views) and the latent gather:
attention over the chunked views must equal full-window attention (fp32 reference SDPA), i.e. concatenating the ranks'
chunks restores the window's key set."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from actionmesh_b200.window_shard import FrameShard, chunked_kv_views, frame_partition, gather_kv

        B, T, L, H, dh = 2, 4, 5, 2, 8
        D = H * dh
        g = torch.Generator().manual_seed(0)  # same on both ranks: the "full window" tensors
        qf = torch.randn(B, T * L, H, dh, generator=g)
        kf = torch.randn(B, T * L, H, dh, generator=g)
        vf = torch.randn(B, T * L, H, dh, generator=g)
        shard = FrameShard()
        fsl = shard.frames(T)
        assert fsl == frame_partition(T, world, rank) and fsl.stop - fsl.start == T // world
        tl = (fsl.stop - fsl.start) * L
        rows = slice(fsl.start * L, fsl.stop * L)
        kv_local = torch.cat([kf[:, rows].reshape(B * tl, D), vf[:, rows].reshape(B * tl, D)], dim=1)  # (B*tl, 2D)
        kv_all = gather_kv(kv_local)
        k5, v5 = chunked_kv_views(kv_all, B, tl, H, dh)
        assert k5.shape == (B, world, tl, H, dh)
        k_cat = k5.reshape(B, world * tl, H, dh)
        v_cat = v5.reshape(B, world * tl, H, dh)
        assert torch.equal(k_cat, kf) and torch.equal(v_cat, vf)  # rank order == frame order
        ql = qf[:, rows]
        ref = torch.nn.functional.scaled_dot_product_attention(qf.transpose(1, 2), kf.transpose(1, 2), vf.transpose(1, 2))
        got = torch.nn.functional.scaled_dot_product_attention(ql.transpose(1, 2), k_cat.transpose(1, 2), v_cat.transpose(1, 2))
        assert torch.allclose(got, ref[:, :, rows], atol=1e-6)
        lat = torch.arange(T * 3 * 2, dtype=torch.float32).reshape(1, T, 3, 2)
        full = shard.gather_latents(lat[:, fsl].contiguous())
        assert torch.equal(full, lat)
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_frame_shard_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_frame_partition_rejects_uneven():
    import pytest

    from actionmesh_b200.window_shard import frame_partition

    assert frame_partition(16, 8, 3) == slice(6, 8)
    with pytest.raises(ValueError):
        frame_partition(16, 3, 0)
