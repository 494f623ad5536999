"""NVLink peer-memory exchange for SyncBatchNorm (csrc/peer.cu, C ABI mvf_peer_allreduce_f64): every rank allocates one symmetric
buffer (torch.distributed._symmetric_memory: CUDA VMM allocations mapped into every rank of the node), the ranks' base pointers go
into a small device table, and each exchange is one single-CTA kernel on the calling stream -- no NCCL call, no host round trip.
`MVF_SYNCBN_PEER=0` keeps the NCCL all-reduce (A/B timing, or nodes without peer access)."""
import os

import torch

from . import _lib

_state = {}      # id(group) -> PeerExchange or False (unavailable)
enabled = os.environ.get("MVF_SYNCBN_PEER", "1") != "0"
exchanges = 0


class PeerExchange:
    def __init__(self, group, device):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        L = _lib.lib()
        self.group, self.device = group, device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        nbytes = L.mvf_peer_buffer_bytes()
        self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, group.group_name if group is not None else dist.group.WORLD.group_name)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        assert len(ptrs) == self.world
        self.ptrs = torch.tensor(ptrs, dtype=torch.int64).to(device)
        self.seq = torch.zeros(8, dtype=torch.int64, device=device)      # one exchange counter per channel
        self.channels = {}                                               # stream handle -> channel
        torch.cuda.synchronize(device)
        dist.barrier(group)                                              # every buffer is zeroed before anyone pushes

    def channel_of(self, stream):
        ch = self.channels.get(stream.cuda_stream)
        if ch is None:
            ch = len(self.channels)   # streams are first used in program order, the same on every rank
            if ch >= 8:
                raise RuntimeError("peer exchange: more than 8 concurrent streams issue SyncBatchNorm exchanges")
            self.channels[stream.cuda_stream] = ch
        return ch

    def allreduce_(self, vec, stream):
        """vec: contiguous float64 device vector, summed over the ranks in place, on `stream`"""
        global exchanges
        ch = self.channel_of(stream)
        _lib.check(_lib.lib().mvf_peer_allreduce_f64(vec.data_ptr(), vec.numel(), self.ptrs.data_ptr(), self.rank, self.world, ch,
                                                     self.seq.data_ptr() + 8 * ch, stream.cuda_stream), "mvf_peer_allreduce_f64")
        exchanges += 1
        return vec


def get(group, device):
    """the exchange object of a process group, or None when peer memory is not usable (then the caller keeps NCCL)"""
    if not enabled:
        return None
    key = id(group)
    st = _state.get(key)
    if st is None:
        try:
            st = PeerExchange(group, device)
        except Exception as e:   # no symmetric memory on this node / group spans nodes
            import sys
            sys.stderr.write("peer exchange unavailable (%s): SyncBatchNorm statistics go through NCCL\n" % str(e).splitlines()[0][:200])
            st = False
        _state[key] = st
    return st or None
