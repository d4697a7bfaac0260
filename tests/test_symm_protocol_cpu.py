"""Host-side logic of the NVLS symmetric-memory exchange (``ops/symm.py::get_vmm_workspace``) against a stand-in extension:
the phases run in order, and a multicast failure on ANY rank (create on rank 0, import/add-device on a peer, bind) makes EVERY rank
report a unicast-only workspace — the kernels on all ranks must agree on the path they take."""
import pytest
import torch

from dist_utils import run_distributed


class _FakeExt:
    def __init__(self, rank, fail):
        self.rank, self.fail, self.calls = rank, fail, []

    def vmm_begin(self, nbytes, rank, world, want_mc):
        self.calls.append("begin")
        return 1 << 20, f"sock{rank}", bool(want_mc) and self.fail != ("nosupport", rank), 1 << 21

    def vmm_send(self, handle, names, use_mc):
        self.calls.append(("send", use_mc))
        assert len(names) == 2 and names[self.rank] == f"sock{self.rank}"
        return "cuMulticastCreate: boom" if (use_mc and self.fail == ("create", self.rank)) else ""

    def vmm_recv(self, handle):
        self.calls.append("recv")
        return "cuMulticastAddDevice: boom" if self.fail == ("recv", self.rank) else ""

    def vmm_bind(self, handle, everyone_ok):
        self.calls.append(("bind", everyone_ok))
        if not everyone_ok:
            return ""
        return "cuMulticastBindMem: boom" if self.fail == ("bind", self.rank) else ""

    def vmm_ptrs(self, handle, mc_everywhere):
        self.calls.append(("ptrs", mc_everywhere))
        return [0x1000, 0x2000], (0x9000 if mc_everywhere else 0), 1 << 21

    def vmm_free(self, handle):
        self.calls.append("free")


def _exchange(rank, world, fail):
    import torch.distributed as dist

    from neuronx_distributed_b200.ops import _ext, symm

    fake = _FakeExt(rank, fail)
    _ext._C, _ext._TRIED = fake, True
    try:
        ws = symm.get_vmm_workspace(dist.group.WORLD, "t", 1 << 20, multicast=True)
        expect_mc = fail is None
        assert ws.has_multicast == expect_mc, (rank, fail, ws.mc_ptr, ws.mc_error)
        assert ws.ptr_list == [0x1000, 0x2000] and ws.local_ptr == [0x1000, 0x2000][rank]
        if fail is not None and fail[0] != "nosupport":
            assert "boom" in ws.mc_error                               # every rank knows why
        if fail == ("nosupport", 1):
            assert ("send", False) in fake.calls                       # multicast is not even attempted
        assert fake.calls[0] == "begin" and fake.calls[-1] == ("ptrs", expect_mc)
        # the same call again is served from the cache; a larger request re-allocates (old region freed)
        assert symm.get_vmm_workspace(dist.group.WORLD, "t", 1 << 20) is ws
        ws2 = symm.get_vmm_workspace(dist.group.WORLD, "t", 4 << 20, multicast=True)
        assert ws2 is not ws and "free" in fake.calls
    finally:
        symm._VMM_WORKSPACES.clear()
        _ext._C, _ext._TRIED = None, False


@pytest.mark.parametrize("fail", [None, ("create", 0), ("recv", 1), ("bind", 1), ("nosupport", 1)])
def test_vmm_exchange_agrees_on_multicast_across_ranks(fail):
    run_distributed(_exchange, 2, fail, timeout=120)
