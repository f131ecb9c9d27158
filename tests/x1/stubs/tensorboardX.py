"""Test stub: tensorboardX is not installed in this image; train_multigpu.py:12 imports SummaryWriter at module level."""


class SummaryWriter:
    def __init__(self, *a, **k):
        pass

    def add_scalar(self, *a, **k):
        pass

    def close(self):
        pass
