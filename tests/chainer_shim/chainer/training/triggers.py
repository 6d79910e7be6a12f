class IntervalTrigger:
    def __init__(self, period, unit):
        self.period, self.unit = period, unit
