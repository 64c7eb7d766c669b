"""Minimal stand-in for mesa 0.8.7 (absent in this image).  The reference only uses these names as base
classes (base_model.py:5-8); Country.__init__ never calls Model.__init__."""


class Agent:
    def __init__(self, unique_id=None, model=None):
        self.unique_id, self.model = unique_id, model


class Model:
    pass
