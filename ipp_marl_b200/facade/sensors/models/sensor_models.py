"""Altitude-dependent noise (reference: sensors/models/sensor_models.py:7-22)."""
from typing import Dict


class AltitudeSensorModel:
    NOISE = {5: 0.01, 10: 0.265, 15: 0.375}

    def __init__(self, params: Dict):
        self.params = params
        self.coeff_a = self.params["sensor"]["model"]["coeff_a"]
        self.coeff_b = self.params["sensor"]["model"]["coeff_b"]

    def get_noise_variance(self, altitude) -> float:
        try:
            return self.NOISE.get(int(altitude), 0) if altitude == int(altitude) else 0
        except (TypeError, ValueError):
            return 0
