"""Host-side logic of x265_b200.lookahead.Lookahead without a device: a fake library records allocations and kernel-call arguments.
Checks that a window's per-step buffers are recycled (a step used to make ~2000 cudaMalloc / cudaFree calls) and that the
result blocks of a batch are laid out as the job records say."""
import numpy as np

from x265_b200.lib import LA_JOB
from x265_b200.lookahead import Lookahead, window_triples, conflict_free_batches


class _Buf:
    def __init__(self, lib, nbytes):
        self.lib, self.nbytes = lib, nbytes
        lib.next_ptr += (nbytes + 255) // 256 * 256 + 256
        self.ptr = lib.next_ptr
        self.data = None

    def upload(self, arr, offset=0):
        self.data = np.ascontiguousarray(arr).copy()
        return self

    def download(self, dtype, count=None, offset=0):
        dt = np.dtype(dtype)
        if count is None:
            count = (self.nbytes - offset) // dt.itemsize
        return np.zeros(count, dt)

    def free(self):
        self.lib.frees += 1
        self.ptr = None


class _Calls:
    def __init__(self, lib):
        self.lib = lib

    def __getattr__(self, name):
        def f(*a):
            self.lib.calls.append((name, a))
            return 0
        return f


class FakeLib:
    def __init__(self):
        self.next_ptr, self.allocs, self.frees, self.calls = 1 << 20, 0, 0, []
        self.ctx = 0
        self.L = _Calls(self)

    def alloc(self, nbytes):
        self.allocs += 1
        return _Buf(self, nbytes)

    def to_device(self, arr):
        b = self.alloc(np.asarray(arr).nbytes)
        return b.upload(arr)

    def check(self, rc):
        assert rc == 0

    def sync(self):
        pass

    def mvcost_table(self, lam, rng):
        return np.zeros(2 * rng + 1, np.uint16)


def _step(la, batches, n):
    la.forget_results()
    for f in la.fr:
        f["has_planes"] = True; f["has_intra"] = True
    preps = [la.prepare_batch(b) for b in batches]
    for p in preps:
        la.launch_batch(p)
    for p in preps:
        la.collect_batch(p, full=True)
    return preps


def test_buffers_are_recycled_and_blocks_match_jobs():
    lib = FakeLib()
    n, W, H = 9, 256, 128
    la = Lookahead(lib, W, H, 8, n, lookahead_slices=0)
    triples = window_triples(n, 3)
    batches = conflict_free_batches(triples)
    assert sorted(t for b in batches for t in b) == sorted(triples)
    base = lib.allocs
    preps = _step(la, batches, n)
    first = lib.allocs - base
    assert first > 0 and lib.frees == 0                       # nothing goes back to the driver between steps
    # the job records of a batch point into its three result blocks, one slot per triple
    for p in preps:
        lcB, rsB, outB = p["blocks"]
        jobs = p["d_jobs"].data.view(LA_JOB) if p["d_jobs"].data.dtype != LA_JOB else p["d_jobs"].data
        nt = len(p["todo"])
        assert len(jobs) == nt                                 # no cooperative slices in this window
        for ti in range(nt):
            assert int(jobs[ti]["lowresCosts"]) == lcB.ptr + ti * 2 * la.ncu
            assert int(jobs[ti]["rowSatds"]) == rsB.ptr + ti * 4 * la.h8
            assert int(jobs[ti]["out"]) == outB.ptr + ti * 32
    # every estimate of the window has a result, B costs carry the 100 / 130 bias on a zero cost (= 0)
    for (p0, p1, b) in triples:
        assert (b - p0, p1 - b) in la.fr[b]["res"]
    a1 = lib.allocs
    _step(la, batches, n)
    _step(la, batches, n)
    assert lib.allocs == a1, "steps after the first must not allocate (%d new allocations)" % (lib.allocs - a1)
    la.close()
    assert lib.frees > 0
