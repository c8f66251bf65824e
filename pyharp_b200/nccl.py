"""A raw NCCL communicator (ncclComm_t) for the C-ABI collectives b200_band_allreduce / b200disort_allreduce_band.

pyharp is a C++ project: its host code would create the communicator with ncclCommInitRank and hand it to the
library.  From Python, torch.distributed keeps its communicators private, so this module creates one directly on the
libnccl that torch already loaded (ctypes), using torch.distributed only to pass the ncclUniqueId around.  The
library itself resolves ncclAllReduce from the same loaded libnccl at run time (pyharp_b200/csrc/band.cu)."""
import ctypes as C
import os


class _UniqueId(C.Structure):
    _fields_ = [("internal", C.c_char * 128)]


def _libnccl():
    import torch  # noqa: F401  (libtorch_cuda pulls in its bundled libnccl.so.2)
    sp = os.path.dirname(os.path.dirname(torch.__file__))
    for cand in (os.path.join(sp, "nvidia", "nccl", "lib", "libnccl.so.2"), "libnccl.so.2"):
        try:
            return C.CDLL(cand, mode=C.RTLD_GLOBAL)
        except OSError:
            continue
    raise RuntimeError("libnccl.so.2 not found")


class Communicator:
    """ncclCommInitRank over the ranks of an initialised torch.distributed process group (any backend)."""

    def __init__(self, device_index):
        import torch
        import torch.distributed as dist
        self._lib = L = _libnccl()
        L.ncclGetUniqueId.restype = C.c_int
        L.ncclGetUniqueId.argtypes = [C.POINTER(_UniqueId)]
        L.ncclCommInitRank.restype = C.c_int
        L.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _UniqueId, C.c_int]
        L.ncclCommDestroy.restype = C.c_int
        L.ncclCommDestroy.argtypes = [C.c_void_p]
        L.ncclGetVersion.restype = C.c_int
        L.ncclGetVersion.argtypes = [C.POINTER(C.c_int)]
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        uid = _UniqueId()
        if self.rank == 0:
            rc = L.ncclGetUniqueId(C.byref(uid))
            if rc != 0:
                raise RuntimeError(f"ncclGetUniqueId failed ({rc})")
        box = [C.string_at(C.byref(uid), 128) if self.rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        C.memmove(C.byref(uid), box[0], 128)
        self._comm = C.c_void_p()
        # NCCL announces its version on stdout at the first communicator; callers (bench.py) own stdout: send it to stderr
        import sys
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            with torch.cuda.device(device_index):
                rc = L.ncclCommInitRank(C.byref(self._comm), self.world, uid, self.rank)
        finally:
            os.dup2(saved, 1)
            os.close(saved)
        if rc != 0:
            raise RuntimeError(f"ncclCommInitRank failed ({rc})")
        v = C.c_int(0)
        L.ncclGetVersion(C.byref(v))
        self.version = v.value

    @property
    def handle(self):
        """ncclComm_t as an integer (pass as void* to the C ABI)."""
        return self._comm.value

    def __int__(self):
        return self._comm.value

    def destroy(self):
        if self._comm:
            self._lib.ncclCommDestroy(self._comm)
            self._comm = C.c_void_p()
