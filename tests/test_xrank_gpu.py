"""The cross-rank batch-norm statistic exchange (csrc/fv_xrank.cu) on ONE GPU: all ranks of a virtual world run as the
blocks of one cooperative launch (fv_bn_finalize_xrank_emulate -- kernels that wait on one another must be co-resident,
so the ranks are not emulated as separate launches), each with its own partial sums, symmetric buffer, epoch counter and
outputs.  Against an fp64 reference of SyncBatchNorm's semantics (torch/nn/modules/_functions.py:39-83, 144-170; reference
modules.py:19): forward = statistics of the union of the ranks' batches; backward = coupling coefficients from the global
sums, dgamma / dbeta from the local ones."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from face_vae_b200 import _lib
    _lib.call("fv_device_ok")
    return _lib


class World:
    def __init__(self, lib, world):
        n = int(lib.load().fv_xrank_buffer_floats())
        self.bufs = [torch.zeros(n, device="cuda") for _ in range(world)]
        self.ptrs = torch.tensor([b.data_ptr() for b in self.bufs], dtype=torch.int64, device="cuda")
        self.epoch = torch.zeros(world, dtype=torch.int64, device="cuda")
        self.world = world


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("c", [32, 256, 512])
def test_forward_exchange_matches_global_batch_statistics(lib, world, c):
    torch.manual_seed(world * 1000 + c)
    W = World(lib, world)
    per_rank = 4096
    count = float(per_rank * world)
    # activations with a mean much larger than their spread (round-1 ADVICE: E[x^2] - E[x]^2 cancellation)
    xs = [(torch.randn(per_rank, c, dtype=torch.float64) * 0.5 + 20.0 + r).cuda() for r in range(world)]
    sums = torch.stack([torch.cat([x.sum(0), (x * x).sum(0)]) for x in xs]).float().contiguous()
    gamma = (torch.rand(world, c, device="cuda") + 0.5)
    beta = torch.rand(world, c, device="cuda") - 0.5
    gamma[:] = gamma[0]
    beta[:] = beta[0]
    rm = torch.zeros(world, c, device="cuda")
    rv = torch.ones(world, c, device="cuda")
    out = torch.empty(world, 4, c, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    lib.call("fv_bn_finalize_xrank_emulate", sums.data_ptr(), W.ptrs.data_ptr(), world, W.epoch.data_ptr(), 0, count, gamma.data_ptr(),
             beta.data_ptr(), rm.data_ptr(), rv.data_ptr(), 0.1, 1e-5, out.data_ptr(), None, None, 0, c, s)
    torch.cuda.synchronize()
    tot = sums.double().sum(0)
    mean = tot[:c] / count
    var = tot[c:] / count - mean * mean
    invstd = 1.0 / torch.sqrt(var + 1e-5)
    for r in range(world):
        assert torch.equal(out[r], out[0]), "every rank must derive bitwise identical statistics"
        torch.testing.assert_close(out[r, 0].double(), mean, rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(out[r, 1].double(), invstd, rtol=2e-3, atol=0)      # fp32 sums of x^2 at mean 20: 1e-3-level variance
        torch.testing.assert_close(out[r, 2].double(), gamma[0].double() * invstd, rtol=2e-3, atol=0)
        torch.testing.assert_close(rm[r].double(), 0.1 * mean, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(rv[r].double(), 0.9 + 0.1 * var * count / (count - 1), rtol=2e-3, atol=1e-6)


@pytest.mark.parametrize("world", [2, 8])
def test_backward_exchange_and_slot_ring(lib, world):
    """Backward mode, and 100 back-to-back exchanges through the ring of 8 slots with changing data."""
    c = 64
    W = World(lib, world)
    s = torch.cuda.current_stream().cuda_stream
    count = 1000.0 * world
    coef = torch.empty(world, 2, c, device="cuda")
    dgamma = torch.empty(world, c, device="cuda")
    dbeta = torch.empty(world, c, device="cuda")
    for it in range(100):
        torch.manual_seed(it)
        local = torch.randn(world, 2 * c, device="cuda") * (1.0 + it)
        lib.call("fv_bn_finalize_xrank_emulate", local.data_ptr(), W.ptrs.data_ptr(), world, W.epoch.data_ptr(), 1, count, None, None, None,
                 None, 0.0, 0.0, coef.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), 0, c, s)
        if it % 33 == 0 or it == 99:
            torch.cuda.synchronize()
            tot = local.double().sum(0) / count
            for r in range(world):
                assert torch.equal(coef[r], coef[0])
                torch.testing.assert_close(coef[r].double().reshape(-1), tot, rtol=1e-5, atol=1e-7)
                torch.testing.assert_close(dbeta[r], local[r, :c])           # local sums, not the global ones
                torch.testing.assert_close(dgamma[r], local[r, c:])
    torch.cuda.synchronize()
    assert int(W.epoch[0]) == 100 and bool((W.epoch == 100).all())


@pytest.mark.parametrize("n", [4, 4 * 1000 + 8, 1 << 20])
def test_grad_allreduce_single_rank_degenerates_to_scaling(lib, n):
    """world = 1 through the C-ABI: the pull / sum / push loop and both flag barriers with no peers (the 2-rank behaviour is
    tests/test_ddp_gpu.py::gradar); five back-to-back epochs exercise the ticket / epoch reset."""
    torch.manual_seed(n)
    words = int(lib.load().fv_grad_allreduce_flag_words())
    buf = torch.randn(n, device="cuda")
    ref = buf.clone()
    flags = torch.zeros(words, dtype=torch.int64, device="cuda")
    bufs = torch.tensor([buf.data_ptr()], dtype=torch.int64, device="cuda")
    fl = torch.tensor([flags.data_ptr()], dtype=torch.int64, device="cuda")
    epoch = torch.zeros(1, dtype=torch.int64, device="cuda")
    ticket = torch.zeros(1, dtype=torch.int32, device="cuda")
    for it in range(5):
        lib.call("fv_grad_allreduce", bufs.data_ptr(), fl.data_ptr(), 0, 1, n, epoch.data_ptr(), ticket.data_ptr(), 0.5,
                 torch.cuda.current_stream().cuda_stream)
        ref = ref * 0.5
    torch.cuda.synchronize()
    assert torch.equal(buf, ref)
    assert int(epoch.item()) == 5 and int(ticket.item()) == 0
