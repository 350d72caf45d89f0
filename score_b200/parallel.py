"""Multi-GPU training of the SCoRe hot path: one process per GPU, torch.distributed (NCCL) for the exchange.

Placeholder until the data-parallel step lands (multi-GPU)."""


class DataParallelTrainer:
    def __init__(self, model, world, rank):
        raise NotImplementedError("data-parallel stepping is not implemented yet")
