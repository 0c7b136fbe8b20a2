"""Host-side mirror of ``wt_simulator.core.chemistry`` for the batched pH operator.

``calculate_pH_batch`` is ``AqueousChemistry(BufferSystem(alk, C_T, T)).calculate_pH(guess)``
(chemistry.py:271-330) over P independent buffer systems, one CUDA thread each.  What the
reference raises becomes a status code per solve.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib

PH_OK = 0
PH_DERIVATIVE_TOO_SMALL = 1   # RuntimeError("Derivative too small ...")  chemistry.py:309-312
PH_NOT_CONVERGED = 2          # RuntimeError("pH calculation did not converge ...")  chemistry.py:327-330
PH_TEMPERATURE_RANGE = 3      # ValueError from celsius_to_kelvin in the constructor


def _dev(x, device):
    if torch.is_tensor(x):
        return x.to(device=device, dtype=torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(device)


def calculate_pH_batch(alkalinity, total_carbonate, temperature, initial_guess, device=None):
    """-> (pH [P] float64, iterations [P] int32, status [P] int32), all on the device."""
    _lib.require_device()
    device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    alk, ct = _dev(alkalinity, device), _dev(total_carbonate, device)
    temp, guess = _dev(temperature, device), _dev(initial_guess, device)
    P = alk.numel()
    if not (ct.numel() == temp.numel() == guess.numel() == P):
        raise ValueError("all inputs must have the same length")
    ph = torch.empty(P, dtype=torch.float64, device=device)
    iters = torch.empty(P, dtype=torch.int32, device=device)
    status = torch.empty(P, dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream().cuda_stream
        p = lambda t: C.c_void_p(t.data_ptr())
        rc = _lib.lib().wt_calc_ph(P, p(alk), p(ct), p(temp), p(guess), p(ph), p(iters), p(status),
                                   C.c_void_p(stream))
    _lib.check(rc, "wt_calc_ph")
    return ph, iters, status


@dataclass
class BufferSystem:
    """chemistry.py:54-80"""

    alkalinity: float
    total_carbonate: float
    temperature: float = 20.0

    def validate(self) -> None:
        if self.alkalinity < 0:
            raise ValueError(f"Alkalinity cannot be negative: {self.alkalinity}")
        if self.total_carbonate < 0:
            raise ValueError(f"Total carbonate cannot be negative: {self.total_carbonate}")


class AqueousChemistry:
    """The calculate_pH / add_acid / add_base corner of chemistry.py:83-398, on the GPU."""

    PH_TOLERANCE = 1e-6
    MAX_ITERATIONS = 100

    def __init__(self, buffer_system: BufferSystem):
        buffer_system.validate()
        if buffer_system.temperature < 0 or buffer_system.temperature > 100:
            raise ValueError(f"Temperature {buffer_system.temperature}°C outside liquid water range [0.0, 100.0]°C.")
        self.buffer = buffer_system

    def calculate_pH(self, initial_guess: float = 7.0) -> float:
        b = self.buffer
        ph, it, st = calculate_pH_batch([b.alkalinity], [b.total_carbonate], [b.temperature], [initial_guess])
        st = int(st[0])
        if st == PH_DERIVATIVE_TOO_SMALL:
            raise RuntimeError(f"Derivative too small at pH={float(ph[0]):.3f}, cannot continue")
        if st == PH_NOT_CONVERGED:
            raise RuntimeError(f"pH calculation did not converge after {self.MAX_ITERATIONS} iterations. "
                               f"Final pH={float(ph[0]):.3f}")
        return float(ph[0])

    def add_acid(self, volume_L: float, acid_mol: float, current_pH: float) -> float:
        delta_alk = -(acid_mol / volume_L) * 50000.0  # chemistry.py:352
        nb = BufferSystem(self.buffer.alkalinity + delta_alk, self.buffer.total_carbonate, self.buffer.temperature)
        return AqueousChemistry(nb).calculate_pH(initial_guess=current_pH)

    def add_base(self, volume_L: float, base_mol: float, current_pH: float) -> float:
        delta_alk = (base_mol / volume_L) * 50000.0  # chemistry.py:385
        nb = BufferSystem(self.buffer.alkalinity + delta_alk, self.buffer.total_carbonate, self.buffer.temperature)
        return AqueousChemistry(nb).calculate_pH(initial_guess=current_pH)
