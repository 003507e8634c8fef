class Normal:
    def __init__(self, loc, scale):
        self.loc, self.scale = loc, scale

    @property
    def mean(self):
        return self.loc

    def rsample(self, *a, **k):
        raise RuntimeError("the shimmed path is deterministic (use_mean=True)")
