"""SensorModel base (reference: sensors/models/__init__.py:1-6)."""


class SensorModel:
    def __init__(self):
        pass
