"""Host-side mirror of the reference's tensor-parallel linear layers over ggml blocks.
  * `SumAllReduce`          — REF mistralrs-quant/src/distributed/mod.rs:432-453 (`is_noop`, `sum_all_reduce`)
  * `ReplicatedLayer`       — REF distributed/layers.rs:1779 (full weight on every rank)
  * `ColumnParallelLayer`   — REF distributed/layers.rs:1160-1294, forward :1610-1616: output rows
                              [rank*N/w, (rank+1)*N/w) of the weight AND of the bias; no exchange
  * `RowParallelLayer`      — REF distributed/layers.rs:695-975, forward :965-975: K columns sharded on quant-block
                              boundaries (REF gguf/weight_source.rs:809-818), partial outputs sum-all-reduced, the FULL
                              bias added once, AFTER the reduce
The linear itself is `quant.GgufMatMul` on the shard (the CUDA kernels; no CPU path).  The exchange is
`torch.distributed` all-reduce by default; the decode stack's in-graph peer-memory all-reduce (`model.PeerAllReduce`)
plugs in through the same two-method interface."""
import torch
import torch.distributed as dist

from . import BLOCK_BYTES, BLOCK_ELEMS


class SumAllReduce:
    def __init__(self, group=None, world_size=None, reduce_fn=None):
        """reduce_fn(tensor) -> tensor overrides the exchange (e.g. PeerAllReduce); default: dist.all_reduce(SUM) in place."""
        self.group, self.reduce_fn = group, reduce_fn
        if world_size is None:
            world_size = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.world_size = world_size

    def is_noop(self):
        return self.world_size == 1

    def sum_all_reduce(self, xs: torch.Tensor) -> torch.Tensor:
        if self.is_noop():
            return xs
        if self.reduce_fn is not None:
            return self.reduce_fn(xs)
        dist.all_reduce(xs, op=dist.ReduceOp.SUM, group=self.group)
        return xs


def _blocks_view(data: torch.Tensor, dtype: str, shape):
    n, k = shape
    if k % BLOCK_ELEMS[dtype]:
        raise ValueError(f"K = {k} is not a whole number of {dtype} blocks")
    return data.reshape(n, k // BLOCK_ELEMS[dtype], BLOCK_BYTES[dtype])


def shard_rows(data: torch.Tensor, dtype: str, shape, rank: int, world: int):
    """Column-parallel shard of raw ggml blocks: rows [rank*N/w, (rank+1)*N/w)."""
    n, k = shape
    if n % world:
        raise ValueError(f"column-parallel rows {n} do not divide by the tensor-parallel size {world}")
    per = n // world
    return _blocks_view(data, dtype, shape)[rank * per:(rank + 1) * per].reshape(-1).contiguous(), (per, k)


def shard_k_blocks(data: torch.Tensor, dtype: str, shape, rank: int, world: int):
    """Row-parallel shard of raw ggml blocks: the rank's contiguous run of K blocks of every row."""
    n, k = shape
    nb = k // BLOCK_ELEMS[dtype]
    if nb % world:
        raise ValueError(f"row-parallel K = {k} does not split into {world} shards on {dtype} block boundaries")
    per = nb // world
    return _blocks_view(data, dtype, shape)[:, rank * per:(rank + 1) * per].contiguous().reshape(-1), (n, k // world)


def _default_linear(data, dtype, shape):
    from . import quant
    return quant.GgufMatMul(quant.QTensor(data, dtype, shape))


class ReplicatedLayer:
    def __init__(self, data, dtype, shape, bias=None, linear_factory=_default_linear):
        self.weight, self.bias, self.shape = linear_factory(data, dtype, tuple(shape)), bias, tuple(shape)

    def forward(self, xs):
        y = self.weight.forward(xs)
        return y if self.bias is None else y + self.bias


class ColumnParallelLayer:
    def __init__(self, data, dtype, shape, rank, world, bias=None, linear_factory=_default_linear):
        shard, self.shape = shard_rows(data, dtype, shape, rank, world)
        self.weight = linear_factory(shard, dtype, self.shape)
        per = self.shape[0]
        self.bias = None if bias is None else bias[rank * per:(rank + 1) * per].contiguous()

    def forward(self, xs):
        y = self.weight.forward(xs)
        return y if self.bias is None else y + self.bias


class RowParallelLayer:
    def __init__(self, data, dtype, shape, rank, world, all_reduce: SumAllReduce, bias=None, linear_factory=_default_linear):
        shard, self.shape = shard_k_blocks(data, dtype, shape, rank, world)
        self.weight, self.bias, self.all_reduce = linear_factory(shard, dtype, self.shape), bias, all_reduce
        self.rank, self.world = rank, world

    def input_slice(self, xs):
        """The rank's K columns of a full-width activation (what the preceding column-parallel layer produced locally)."""
        kl = self.shape[1]
        return xs[..., self.rank * kl:(self.rank + 1) * kl]

    def forward(self, xs):
        y = self.weight.forward(xs)
        if not self.all_reduce.is_noop():
            y = self.all_reduce.sum_all_reduce(y.contiguous())
        return y if self.bias is None else y + self.bias      # once, after the reduce — not world_size times
