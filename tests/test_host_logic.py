"""CPU tests of the host-side plumbing that needs no GPU: rank ownership after a rebalance, segment
capacities for moved chunks, and the transport's choice between the synchronous and the
lagged-count particle exchange."""
import numpy as np

from picnix_b200 import capi, distributed


def test_owner_and_capacities():
    boundary = [0, 3, 3 + 4, 16]
    owners = [distributed._owner(boundary, g) for g in range(16)]
    assert owners == [0] * 3 + [1] * 4 + [2] * 9
    caps = distributed._capacities(np.array([[100, 0], [1000, 10]]), 0.2)
    assert caps.dtype == np.int32 and caps.tolist() == [120 + 256, 256, 1200 + 256, 12 + 256]


def test_rebalance_boundaries_are_valid_partitions():
    rng = np.random.default_rng(0)
    for nchunk, nrank in [(16, 3), (64, 8), (512, 8)]:
        load = rng.uniform(0.1, 5.0, nchunk)
        b0 = capi.assign_initial(np.ones(nchunk), nrank)
        for b in (capi.assign_initial(load, nrank), capi.assign_rebalance(load, b0)):
            assert b[0] == 0 and b[-1] == nchunk and np.all(np.diff(b) >= 1)
            # every chunk has exactly one owner
            assert sorted(distributed._owner(b, g) for g in range(nchunk)) == sorted(
                np.repeat(np.arange(nrank), np.diff(b)).tolist())


class _FakeSim:
    """comm_buffer() of an arena: recv size 0 = unknown (synchronous protocol), > 0 = lagged-count."""

    def __init__(self, recv_bytes):
        self.recv_bytes = recv_bytes
        self.set_calls = []

    def peers(self):
        return [1]

    def comm_buffer(self, mode, i):
        return 0, 64, 0, self.recv_bytes

    def set_recv_bytes(self, mode, i, n):
        self.set_calls.append(n)


def test_transport_skips_count_exchange_when_sizes_are_known(monkeypatch):
    calls = []
    for recv_bytes, expect_count_exchange in ((0, True), (128, False)):
        sim = _FakeSim(recv_bytes)
        tr = distributed.Transport(sim, world=2, device_buffers=False)

        def fake_exchange(pairs, _calls=calls):
            _calls.append(len(pairs))
            for _, s, r in pairs:
                if r.numel() == 1 and r.dtype.is_floating_point is False and s.numel() == 1:
                    r.copy_(s)

        monkeypatch.setattr(tr, "_exchange", fake_exchange)
        monkeypatch.setattr(tr, "view", lambda ptr, n: __import__("torch").zeros(n, dtype=__import__("torch").uint8))
        calls.clear()
        tr.move(distributed.MODE_PARTICLE)
        assert len(calls) == (2 if expect_count_exchange else 1)
        assert bool(sim.set_calls) == expect_count_exchange
