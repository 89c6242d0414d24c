"""Host-side partition logic of the sharded path (mirrors shard_range() in csrc/exact_kernels.cu and the row split in
csrc/barrier_kernels.cu): rank r of P owns the contiguous slice [n*r//P, n*(r+1)//P) of query primitives / rows."""


def shard_range(n, rank, nranks):
    return (n * rank) // nranks, (n * (rank + 1)) // nranks


def combine_energy(parts):
    """ncclAllReduce(sum) of the per-rank partial energies"""
    return float(sum(parts))


def combine_step(parts):
    """ncclAllReduce(min) of the per-rank CCD steps"""
    return float(min(parts))
