from collections import OrderedDict


class BaseScheduler:
    def __init__(self, model):
        self.model = model
        self.steps = 0
        self.time = 0
        self._agents = OrderedDict()


class RandomActivation(BaseScheduler):
    pass
