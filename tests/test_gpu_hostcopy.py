"""
The host copy engine (csrc/ht_hostcopy.cu) and the bit-packed way out of the device: ht_bits_to_rowbits,
the pool's upload / download / download+expand jobs, and transform_host's two result paths -- always
bit-exact against the oracle / numpy.
"""
import numpy as np
import pytest
import torch

from oracle import transform_oracle as orc

pytestmark = pytest.mark.gpu

from haptools_b200 import _cuda, engine  # noqa: E402
from haptools_b200 import dist as hdist  # noqa: E402
from test_gpu_parity import _layouts, _random_arrays  # noqa: E402


def _rowbits_want(result, pitch):
    n, H, _ = result.shape
    flat = np.zeros((n, pitch * 8), dtype=np.uint8)
    flat[:, : 2 * H] = result.reshape(n, 2 * H)
    return np.packbits(flat, axis=1, bitorder="little")


@pytest.mark.parametrize("n,H", [(1, 1), (31, 15), (32, 16), (33, 17), (1023, 511), (1024, 512), (1025, 513), (2504, 300), (4100, 1030)])
def test_bits_to_rowbits(n, H):
    lib = _cuda.lib()
    rng = np.random.default_rng(n * 7 + H)
    result = rng.random((n, H, 2)) < 0.4
    bits = torch.from_numpy(hdist.pack_result_bits(result).view(np.int32)).cuda()
    for extra in (0, 16):
        pitch = int(lib.ht_rowbits_pitch(H)) + extra
        rb = torch.full((n + 3, pitch), 0xEE, dtype=torch.uint8, device="cuda")
        _cuda.check(lib.ht_bits_to_rowbits(bits.data_ptr(), n, H, rb.data_ptr(), pitch, 0), "ht_bits_to_rowbits")
        got = rb.cpu().numpy()
        used = (H + 15) // 16 * 4
        np.testing.assert_array_equal(got[:n, :used], _rowbits_want(result, pitch)[:, :used])
        assert (got[n:] == 0xEE).all() and (got[:n, used:] == 0xEE).all()


@pytest.mark.parametrize("pinned", [False, True])
def test_download_expand_job(pinned):
    lib = _cuda.lib()
    rng = np.random.default_rng(5)
    n, H = 5000, 40_000  # 400 MB of result: many blocks per worker
    width = (2 * H + 7) // 8
    pitch = int(lib.ht_rowbits_pitch(H))
    rowbits = rng.integers(0, 256, size=(n, pitch), dtype=np.uint8)
    dev = torch.from_numpy(rowbits).cuda()
    out_pitch = 2 * H + 32
    host = engine.pinned_empty((n, out_pitch)) if pinned else np.empty((n, out_pitch), dtype=np.uint8)
    host[:] = 0x77
    t = lib.ht_host_submit_download_expand(host.ctypes.data, out_pitch, 2 * H, dev.data_ptr(), pitch, n, 0)
    assert t > 0
    _cuda.check(lib.ht_host_wait(t), "wait")
    _cuda.check(lib.ht_host_wait(t), "wait twice is harmless")
    want = np.unpackbits(rowbits[:, :width], axis=1, bitorder="little")[:, : 2 * H]
    np.testing.assert_array_equal(host[:, : 2 * H], want)
    assert (host[:, 2 * H:] == 0x77).all()


def test_upload_and_download_jobs_many_outstanding():
    lib = _cuda.lib()
    assert lib.ht_host_workers() >= 1
    rng = np.random.default_rng(6)
    jobs = []
    for i in range(6):
        rows, width = int(rng.integers(200, 3000)), int(rng.integers(1000, 40_000))
        src = rng.integers(0, 256, size=(rows, width + 24), dtype=np.uint8)
        idx = np.sort(rng.choice(rows, size=rows // 3, replace=False)).astype(np.uint32) if i % 2 else None
        n_out = len(idx) if idx is not None else rows
        dpitch = (width + 15) // 16 * 16
        dbuf = torch.zeros((n_out, dpitch), dtype=torch.uint8, device="cuda")
        t = lib.ht_host_submit_upload(dbuf.data_ptr(), dpitch, src.ctypes.data, src.strides[0], width, n_out,
                                      idx.ctypes.data if idx is not None else None, 0)
        assert t > 0
        jobs.append((t, src, idx, width, dbuf))
    _cuda.check(lib.ht_host_wait(0), "wait for all")
    back = []
    for t, src, idx, width, dbuf in jobs:
        want = src[idx if idx is not None else slice(None), :width]
        np.testing.assert_array_equal(dbuf[:, :width].cpu().numpy(), want)
        host = np.zeros((dbuf.shape[0], width + 8), dtype=np.uint8)
        back.append((lib.ht_host_submit_download(host.ctypes.data, width + 8, dbuf.data_ptr(), dbuf.stride(0), width,
                                                 dbuf.shape[0], 0), host, want))
    for t, host, want in back:
        assert t > 0
        _cuda.check(lib.ht_host_wait(t), "wait")
        np.testing.assert_array_equal(host[:, : want.shape[1]], want)
        assert (host[:, want.shape[1]:] == 0).all()


def test_job_waits_for_the_callers_stream():
    """The job's precondition is what the caller queued on its stream before submitting."""
    lib = _cuda.lib()
    n, w = 4096, 4096
    a = torch.zeros((n, w), dtype=torch.uint8, device="cuda")
    st = torch.cuda.Stream()
    host = np.zeros((n, w), dtype=np.uint8)
    with torch.cuda.stream(st):
        for _ in range(200):  # a long queue of work that ends with the value the download must see
            a.add_(1)
        t = lib.ht_host_submit_download(host.ctypes.data, w, a.data_ptr(), w, w, n, st.cuda_stream)
    _cuda.check(lib.ht_host_wait(t), "wait")
    assert (host == 200).all()


def test_submit_rejects_bad_arguments():
    lib = _cuda.lib()
    a = torch.zeros(64, dtype=torch.uint8, device="cuda")
    h = np.zeros(64, dtype=np.uint8)
    assert lib.ht_host_submit_upload(a.data_ptr(), 8, h.ctypes.data, 4, 8, 4, None, 0) == -1  # pitch < width
    assert b"pitch" in lib.ht_last_error()
    assert lib.ht_host_submit_download(None, 8, a.data_ptr(), 8, 8, 4, 0) == -1
    assert lib.ht_host_submit_download_expand(h.ctypes.data, 8, 16, a.data_ptr(), 2, 4, 0) == -1  # out pitch < row
    big = 9 << 20
    assert lib.ht_host_submit_upload(a.data_ptr(), big, h.ctypes.data, big, big, 1, None, 0) == -1  # row > slot
    assert lib.ht_host_wait(-5) == _cuda.HT_ERR_BAD_ARG


@pytest.mark.parametrize("result", ["u8", "bits"])
@pytest.mark.parametrize("seed", range(12))
def test_fuzz_result_paths_vs_oracle(seed, result):
    rng = np.random.default_rng(7000 + seed)
    n = int(rng.choice([1, 33, 1024, 1500, 2100, 4097]))
    p = int(rng.integers(1, 90))
    H = int(rng.choice([1, 3, 17, 128, 130, 400, 513]))
    anc_kind = int(rng.integers(0, 3))
    gt, anc, plan, want = _random_arrays(
        rng, n, p, H, max_k=int(rng.integers(1, 8)), multi=bool(rng.integers(0, 2)),
        missing=float(rng.choice([0.0, 0.01, 0.2])), n_pops=[0, 5, 20][anc_kind], long_hap=int(rng.choice([0, 60])),
    )
    lay = str(rng.choice(["C", "F", "C3", "F3"]))
    arr = _layouts(gt)[lay]
    a = _layouts(anc)[lay] if anc is not None else None
    chunk = int(rng.choice([0, 1024, 2048])) or None
    out = None
    if rng.integers(0, 2):
        out = engine.pinned_empty((n, H, 2))
    got = engine.transform_host(arr, plan, ancestry=a, chunk_samples=chunk, result=result, out=out)
    assert got.dtype == np.bool_ and got.flags.c_contiguous
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("result", ["u8", "bits"])
@pytest.mark.parametrize("lay", ["C", "F"])
def test_large_pageable_arrays_go_through_the_pool(lay, result):
    """Source and result of tens of MB in ordinary numpy memory: uploads and downloads are host copy engine
    jobs (several chunks in flight), and the result is what the oracle says."""
    rng = np.random.default_rng(21)
    n, p, H = 12_000, 700, 900  # 16.8 MB of genotypes, 21.6 MB of result
    gt, _, plan, want = _random_arrays(rng, n, p, H, max_k=4, missing=0.01)
    arr = _layouts(gt)[lay]
    assert _cuda.lib().ht_host_is_pageable(arr.ctypes.data) == 1
    got = engine.transform_host(arr, plan, chunk_samples=3072, result=result)
    np.testing.assert_array_equal(got, want)
    pinned_in = engine.pinned_empty(gt.shape)
    pinned_in[:] = gt
    assert _cuda.lib().ht_host_is_pageable(pinned_in.ctypes.data) == 0
    got = engine.transform_host(pinned_in, plan, chunk_samples=3072, result=result, out=engine.pinned_empty((n, H, 2)))
    np.testing.assert_array_equal(got, want)


def test_auto_picks_bits_for_large_results(monkeypatch):
    rng = np.random.default_rng(22)
    n, p, H = 6000, 64, 800  # 9.6 MB of result >= RESULT_BITS_MIN
    gt, _, plan, want = _random_arrays(rng, n, p, H, max_k=3)
    seen = []
    real = engine.transform_bits

    def spy(*a, **kw):
        seen.append(1)
        return real(*a, **kw)

    monkeypatch.setattr(engine, "transform_bits", spy)
    np.testing.assert_array_equal(engine.transform_host(gt, plan), want)
    assert seen, "a result of 9.6 MB should have left the device bit-packed"
    seen.clear()
    monkeypatch.setenv("HAPTOOLS_B200_RESULT", "u8")
    np.testing.assert_array_equal(engine.transform_host(gt, plan), want)
    assert not seen


@pytest.mark.parametrize("result", ["u8", "bits"])
@pytest.mark.parametrize("depth", ["2", "3", ""])
def test_pipe_depth_and_on_queued(monkeypatch, depth, result):
    """Buffer sets per stage of transform_host's pipeline: 2 (what the device needs; buffers are reused and the
    queueing loop waits), 3, and the default (a set per chunk when device memory allows: nothing is reused and the
    whole call is queued before the first wait) -- same result, and `on_queued` fires exactly once, with a
    page-locked and with a pageable source."""
    if depth:
        monkeypatch.setenv("HAPTOOLS_B200_PIPE_DEPTH", depth)
    else:
        monkeypatch.delenv("HAPTOOLS_B200_PIPE_DEPTH", raising=False)
    rng = np.random.default_rng(33)
    n, p, H = 9_000, 700, 650
    gt, anc, plan, want = _random_arrays(rng, n, p, H, max_k=4, missing=0.01, n_pops=5)
    pinned_in, pinned_anc = engine.pinned_empty(gt.shape), engine.pinned_empty(anc.shape)
    pinned_in[:], pinned_anc[:] = gt, anc
    for g, a in ((pinned_in, pinned_anc), (gt, anc)):
        fired = []
        got = engine.transform_host(g, plan, ancestry=a, chunk_samples=1024, result=result, on_queued=lambda: fired.append(1))
        np.testing.assert_array_equal(got, want)
        assert fired == [1]
