import os

SUM = "sum"


class _Comm:
    def __init__(self):
        self._rank = int(os.environ.get("RANK", os.environ.get("OMPI_COMM_WORLD_RANK", "0")))
        self._size = int(os.environ.get("WORLD_SIZE", os.environ.get("OMPI_COMM_WORLD_SIZE", "1")))

    def _dist(self):
        import torch.distributed as dist
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29541")
            dist.init_process_group("gloo", rank=self._rank, world_size=self._size)
        return dist

    def Get_rank(self): return self._rank
    def Get_size(self): return self._size
    rank = property(Get_rank)
    size = property(Get_size)

    def barrier(self):
        if self._size > 1:
            self._dist().barrier()
    Barrier = barrier

    def bcast(self, obj, root=0):
        if self._size == 1:
            return obj
        box = [obj]
        self._dist().broadcast_object_list(box, src=root)
        return box[0]

    def gather(self, obj, root=0):
        if self._size == 1:
            return [obj]
        out = [None] * self._size if self._rank == root else None
        self._dist().gather_object(obj, out, dst=root)
        return out

    def allgather(self, obj):
        if self._size == 1:
            return [obj]
        out = [None] * self._size
        self._dist().all_gather_object(out, obj)
        return out

    def reduce(self, obj, op=SUM, root=0):
        """element-wise sum of numbers / numpy arrays on `root` (None elsewhere)"""
        parts = self.gather(obj, root=root)
        if parts is None:
            return None
        total = parts[0]
        for p in parts[1:]:
            total = total + p
        return total


COMM_WORLD = _Comm()
