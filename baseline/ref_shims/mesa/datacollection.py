class DataCollector:
    def __init__(self, *a, **k):
        pass
