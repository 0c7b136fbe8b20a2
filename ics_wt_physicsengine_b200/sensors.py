"""Host-side mirror of ``wt_simulator.sensors`` for the batched sensor suite.

``create_realistic_sensor_suite(ensemble)`` builds, for every plant of a ``PlantEnsemble``, the 7
sensors of the reference factory (sensors/__init__.py:41-120) and ``SensorSuite.read(state, t)`` is
one kernel launch doing ``<Sensor>.read(reactor_state, current_time)`` for all of them in the
reference's dict order (base_sensor.py:509-699 + the four subclasses).  Enum values, reading
fields, calibration semantics (``calibrate(reference, t)`` -> offset = reference - current_value)
and the monotonic-time ``ValueError`` are the reference's.  Randomness is a counter-based Philox
stream keyed by (seed, global plant id, read index): outputs match the reference in
distribution, not draw for draw (the reference seeds from ``secrets``, base_sensor.py:331).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from enum import Enum
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from .ensembles import CFG_FIELDS


class SensorStatus(Enum):  # base_sensor.py:49-63
    NORMAL = "normal"
    CALIBRATING = "calibrating"
    WARMING_UP = "warming_up"
    FAILED = "failed"
    SATURATED = "saturated"
    DRIFT_WARNING = "drift_warning"
    CALIBRATION_EXPIRED = "calibration_expired"
    OPEN_CIRCUIT = "open_circuit"
    SHORT_CIRCUIT = "short_circuit"
    OUT_OF_RANGE = "out_of_range"
    POWER_FAULT = "power_fault"
    RATE_OF_CHANGE_FAULT = "rate_of_change_fault"


class SensorFault(Enum):  # base_sensor.py:66-75
    NONE = "none"
    OPEN_CIRCUIT = "open_circuit"
    SHORT_CIRCUIT = "short_circuit"
    OUT_OF_RANGE = "out_of_range"
    RATE_FAULT = "rate_fault"
    POWER_LOW = "power_low"
    POWER_HIGH = "power_high"


class TemperatureSensorType(Enum):  # temperature_sensor.py:29-35
    RTD_PT100 = "rtd_pt100"
    RTD_PT1000 = "rtd_pt1000"
    THERMOCOUPLE_K = "thermocouple_k"
    THERMOCOUPLE_J = "thermocouple_j"


class FlowSensorType(Enum):  # flow_sensor.py:33-37
    TURBINE = "turbine"
    MAGNETIC = "magnetic"


_TEMP_CODE = {TemperatureSensorType.RTD_PT100: 0, TemperatureSensorType.RTD_PT1000: 1,
              TemperatureSensorType.THERMOCOUPLE_K: 2, TemperatureSensorType.THERMOCOUPLE_J: 3}
_FLOW_CODE = {FlowSensorType.MAGNETIC: 0, FlowSensorType.TURBINE: 1}

STATUS_BY_CODE = tuple(SensorStatus)   # device codes are the enum declaration order
FAULT_BY_CODE = tuple(SensorFault)
SENSOR_NAMES = ("pH_inlet", "pH_outlet", "chlorine_inlet", "chlorine_outlet", "flow_main", "temp_inlet", "temp_outlet")


@dataclass
class InstallationQuality:  # base_sensor.py:124-145; defaults = the suite's "good_installation"
    flow_velocity: float = 0.5
    air_bubble_frequency: float = 0.0
    grounding_quality: float = 0.9
    pipe_vibration_g: float = 0.1
    ambient_temperature: float = 30.0

    def validate(self):
        if not 0.0 <= self.flow_velocity <= 5.0:
            raise ValueError(f"Flow velocity {self.flow_velocity} m/s out of range")
        if not 0.0 <= self.grounding_quality <= 1.0:
            raise ValueError("Grounding quality must be 0-1")
        if self.pipe_vibration_g < 0:
            raise ValueError("Vibration must be non-negative")


@dataclass
class SampleLine:  # base_sensor.py:148-175; defaults = the suite's sample lines
    volume_mL: float = 250.0
    flow_rate_mL_min: float = 500.0
    ambient_temp: float = 25.0

    @property
    def transport_delay_s(self) -> float:
        volume_L = self.volume_mL / 1000.0
        flow_rate_L_s = self.flow_rate_mL_min / 1000.0 / 60.0
        return volume_L / flow_rate_L_s if flow_rate_L_s > 0 else 0.0


@dataclass
class BatchReading:
    """SensorReading (base_sensor.py:78-103) for P plants: device tensors of length P."""

    timestamp: float
    value: torch.Tensor
    raw_value: torch.Tensor
    noise: torch.Tensor
    drift: torch.Tensor
    status: torch.Tensor       # int32 codes, see STATUS_BY_CODE
    uncertainty: torch.Tensor
    fault: torch.Tensor        # int32 codes, see FAULT_BY_CODE

    def status_of(self, p: int) -> SensorStatus:
        return STATUS_BY_CODE[int(self.status[p])]

    def fault_of(self, p: int) -> SensorFault:
        return FAULT_BY_CODE[int(self.fault[p])]


class SensorSuite:
    """The 7-sensor suite of every plant of one ensemble shard, state resident in HBM."""

    def __init__(self, ensemble, seed: int = 0, plant0: int = 0, installation: Optional[InstallationQuality] = None,
                 sample_line: Optional[SampleLine] = None, history: int = 0,
                 temperature_sensor_type: TemperatureSensorType = TemperatureSensorType.RTD_PT100,
                 flow_sensor_type: FlowSensorType = FlowSensorType.MAGNETIC):
        _lib.require_device()
        self.ens = ensemble
        self.seed, self.plant0 = int(seed) & (2 ** 64 - 1), int(plant0)
        self.installation = installation or InstallationQuality()
        self.installation.validate()
        line = sample_line or SampleLine()
        if int(line.transport_delay_s) + 10 > 100:
            raise ValueError("sample-line delay needs a deque longer than the 100 slots the engine keeps")
        i = self.installation
        # the factory builds RTD PT100 and magnetic sensors (sensors/__init__.py:100-118); the other variants of
        # TemperatureSensor / FlowSensor are selected for the whole suite
        self.temperature_sensor_type = TemperatureSensorType(temperature_sensor_type)
        self.flow_sensor_type = FlowSensorType(flow_sensor_type)
        self._suite6 = np.array([i.flow_velocity, i.air_bubble_frequency, i.grounding_quality, i.pipe_vibration_g,
                                 i.ambient_temperature, line.transport_delay_s,
                                 _TEMP_CODE[self.temperature_sensor_type], _FLOW_CODE[self.flow_sensor_type]], dtype=np.float64)
        P, dev = ensemble.n_plants, ensemble.device
        cfg = ensemble.cfg
        col = lambda k: torch.from_numpy(np.ascontiguousarray(cfg[:, CFG_FIELDS.index(k)])).to(dev)
        self._cfg_flow, self._cfg_cl, self._cfg_T = col("flow_rate"), col("initial_chlorine"), col("temperature")
        f64, i32 = torch.float64, torch.int32
        self._sens = torch.zeros((10, 7, P), dtype=f64, device=dev)   # WT_NSF fields
        self._sens_i = torch.zeros((4, 7, P), dtype=i32, device=dev)   # WT_NSI: status, fault, len(history), flags
        self._clock = None   # device clock {t, t_prev, read_index, t0, dt} of captured (CUDA graph) reads
        self._ring = torch.zeros((2, 100, 2, P), dtype=f64, device=dev)
        self._ring_i = torch.zeros((2, 2, P), dtype=i32, device=dev)
        self._out = torch.zeros((5, 7, P), dtype=f64, device=dev)
        self._out_status = torch.zeros((7, P), dtype=i32, device=dev)
        self._out_fault = torch.zeros((7, P), dtype=i32, device=dev)
        self.read_index = 0
        self.last_time: Optional[float] = None
        self._initialized = False
        # optional reading history for get_statistics (the reference keeps 1000 readings per sensor,
        # base_sensor.py:248, 321; here a ring of `history` reads x 7 values per plant, off by default)
        self.history = int(history)
        self._hist = torch.zeros((self.history, 7, P), dtype=f64, device=dev) if self.history > 0 else None
        self._hist_times: list = []

    def keys(self):
        return SENSOR_NAMES

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def initialize(self, t0: float) -> None:
        """Constructors + __main__.initialize_sensors: calibrate every sensor at t0 with references
        7.0 (pH), config.initial_chlorine, config.temperature, config.flow_rate (__main__.py:96-105)."""
        p = lambda t: C.c_void_p(t.data_ptr())
        with torch.cuda.device(self.ens.device):
            rc = _lib.lib().wt_sensors_init(self.ens.n_plants, float(t0), p(self._cfg_flow), p(self._cfg_cl), p(self._cfg_T),
                                            p(self._sens), p(self._sens_i), p(self._ring_i), self._stream())
        _lib.check(rc, "wt_sensors_init")
        self.read_index, self.last_time, self._initialized = 0, None, True
        self._t0 = float(t0)

    def calibrate(self, name: str, reference, current_time: float) -> None:
        """BaseSensor.calibrate(reference, t) for sensor `name` of every plant (base_sensor.py:701-755)."""
        s = SENSOR_NAMES.index(name)
        ref_t, ref_s = None, 0.0
        if np.ndim(reference) == 0 and not torch.is_tensor(reference):
            ref_s = float(reference)
        else:
            ref_t = torch.as_tensor(reference, dtype=torch.float64).reshape(self.ens.n_plants).to(self.ens.device).contiguous()
        with torch.cuda.device(self.ens.device):
            rc = _lib.lib().wt_sensors_calibrate(self.ens.n_plants, s, float(current_time),
                                                 C.c_void_p(0 if ref_t is None else ref_t.data_ptr()), ref_s,
                                                 C.c_void_p(self._sens.data_ptr()), C.c_void_p(self._sens_i.data_ptr()),
                                                 self._stream())
        _lib.check(rc, "wt_sensors_calibrate")

    def read(self, reactor_state=None, current_time: Optional[float] = None) -> Dict[str, BatchReading]:
        """All 7 sensors of every plant read the ensemble state at `current_time`."""
        if not self._initialized:
            raise RuntimeError("call initialize(t0) first (the reference calibrates its sensors at start-up)")
        if current_time is None:
            raise ValueError("current_time is required (the ensemble runs on simulated time)")
        if self.last_time is not None and current_time < self.last_time:
            raise ValueError(f"Non-monotonic time: {current_time} < {self.last_time}")  # base_sensor.py:543-549
        e = self.ens
        p = lambda t: C.c_void_p(t.data_ptr())
        t_prev = self.last_time if self.last_time is not None else float(current_time)
        with torch.cuda.device(e.device):
            rc = _lib.lib().wt_sensors_read(
                e.n_plants, e.n_zones, self.plant0, self.read_index, float(current_time), float(t_prev), p(e._y), p(e._flow),
                p(self._cfg_flow), p(self._cfg_cl), p(self._cfg_T), p(self._sens), p(self._sens_i), p(self._ring),
                p(self._ring_i), p(self._out), p(self._out_status), p(self._out_fault),
                self._suite6.ctypes.data_as(C.POINTER(C.c_double)), C.c_uint64(self.seed), C.c_void_p(0), self._stream())
        _lib.check(rc, "wt_sensors_read")
        if self._hist is not None:
            self._hist[self.read_index % self.history].copy_(self._out[0])
            self._hist_times.append(float(current_time))
            if len(self._hist_times) > self.history:
                self._hist_times.pop(0)
        self.read_index += 1
        self.last_time = float(current_time)
        o = self._out
        return {name: BatchReading(float(current_time), o[0, s], o[1, s], o[2, s], o[3, s], self._out_status[s], o[4, s],
                                   self._out_fault[s]) for s, name in enumerate(SENSOR_NAMES)}


    # ---- captured reads: the time of the read lives on the device so that a CUDA graph can be replayed ----
    def start_clock(self, t_next: float, dt: float) -> None:
        """Device clock for ``read_clocked``: the next read happens at ``t_next``, every following one ``dt`` later."""
        if not self._initialized:
            raise RuntimeError("call initialize(t0) first")
        k = float(self.read_index)
        t_prev = self.last_time if self.last_time is not None else float(t_next)
        self._clock_host = (float(t_next), float(dt))
        row = torch.tensor([float(t_next), float(t_prev), k, float(t_next) - k * float(dt), float(dt)], dtype=torch.float64)
        if self._clock is None:
            self._clock = torch.empty(5, dtype=torch.float64, device=self.ens.device)
        self._clock.copy_(row)

    def read_clocked(self) -> None:
        """One suite read at the device clock's time followed by a clock tick (two launches, no host scalars:
        capturable).  The caller accounts for the reads with ``account_reads`` after the (re)play."""
        e = self.ens
        p = lambda t: C.c_void_p(t.data_ptr())
        with torch.cuda.device(e.device):
            L = _lib.lib()
            rc = L.wt_sensors_read(
                e.n_plants, e.n_zones, self.plant0, 0, 0.0, 0.0, p(e._y), p(e._flow),
                p(self._cfg_flow), p(self._cfg_cl), p(self._cfg_T), p(self._sens), p(self._sens_i), p(self._ring),
                p(self._ring_i), p(self._out), p(self._out_status), p(self._out_fault),
                self._suite6.ctypes.data_as(C.POINTER(C.c_double)), C.c_uint64(self.seed), p(self._clock), self._stream())
            _lib.check(rc, "wt_sensors_read")
            _lib.check(L.wt_clock_tick(p(self._clock), self._stream()), "wt_clock_tick")

    def account_reads(self, n_reads: int) -> None:
        """Host-side bookkeeping of ``n_reads`` clocked reads (read index, last read time)."""
        t_next, dt = self._clock_host
        self.read_index += int(n_reads)
        self.last_time = t_next + (int(n_reads) - 1) * dt
        self._clock_host = (t_next + int(n_reads) * dt, dt)

    def readings(self) -> Dict[str, BatchReading]:
        """The readings of the last read (views of the device output buffers)."""
        o = self._out
        t = self.last_time if self.last_time is not None else float("nan")
        return {name: BatchReading(t, o[0, s], o[1, s], o[2, s], o[3, s], self._out_status[s], o[4, s], self._out_fault[s])
                for s, name in enumerate(SENSOR_NAMES)}

    def reset(self, name: str, current_time: float) -> None:
        """BaseSensor.reset() (base_sensor.py:858-878) for sensor `name` of every plant; the reference stamps
        time.monotonic() into the calibration / power-on times, the ensemble runs on simulated time, so the time is
        an argument.  Every later read reports CALIBRATION_EXPIRED until ``calibrate`` (the reference clears its
        calibration history, base_sensor.py:432-436, 870)."""
        s = SENSOR_NAMES.index(name)
        p = lambda t: C.c_void_p(t.data_ptr())
        with torch.cuda.device(self.ens.device):
            rc = _lib.lib().wt_sensors_reset(self.ens.n_plants, s, float(current_time), p(self._cfg_flow), p(self._sens),
                                             p(self._sens_i), p(self._ring_i), self._stream())
        _lib.check(rc, "wt_sensors_reset")

    # ---- maintenance: the reference's per-sensor methods, for that sensor of every plant ----
    def _maintain(self, name: str, op: int, t: float, a0: float = 0.0, a1: float = 0.0) -> None:
        s = SENSOR_NAMES.index(name)
        with torch.cuda.device(self.ens.device):
            rc = _lib.lib().wt_sensors_maintain(self.ens.n_plants, s, op, float(t), float(a0), float(a1),
                                                C.c_void_p(self._sens.data_ptr()), C.c_void_p(self._sens_i.data_ptr()),
                                                self._stream())
        if rc == -1:   # WT_ERR_BAD_ARG: the reference raises ValueError / AttributeError for these
            raise ValueError(_lib.lib().wt_last_error().decode(errors="replace"))
        _lib.check(rc, "wt_sensors_maintain")

    def calibrate_two_point(self, name: str, buffer_pH_1: float, buffer_pH_2: float, measured_pH_1: float,
                            measured_pH_2: float, current_time: float) -> None:
        """pHSensor.calibrate_two_point (ph_sensor.py:338-393).  The measured values only set slope_percentage,
        which the next read() overwrites (ph_sensor.py:256-262)."""
        self._maintain(name, 0, current_time, buffer_pH_1, buffer_pH_2)

    def clean_electrode(self, name: str, cleaning_method: str, current_time: float) -> None:
        """pHSensor.clean_electrode (ph_sensor.py:395-434)."""
        methods = {"water_rinse": 0, "acid_clean": 1, "pepsin_clean": 2}
        if cleaning_method not in methods:
            raise ValueError(f"Unknown cleaning method: {cleaning_method}")
        self._maintain(name, 1, current_time, methods[cleaning_method])

    def replace_membrane(self, name: str, current_time: float) -> None:
        """ChlorineSensor.replace_membrane (chlorine_sensor.py:486-509): amperometric sensors only."""
        self._maintain(name, 2, current_time)

    def replace_reagent(self, name: str, current_time: float) -> None:
        """ChlorineSensor.replace_reagent (chlorine_sensor.py:511-537): DPD sensors only."""
        self._maintain(name, 3, current_time)

    def get_statistics(self, name: str, window_seconds: float = 60.0) -> Dict[str, torch.Tensor]:
        """BaseSensor.get_statistics (base_sensor.py:809-856) of sensor `name` for every plant, over the readings
        of the last `window_seconds` that are still in the history ring (construct the suite with history=K).
        Returns {mean, std, min, max, count, drift_rate, fault_rate}: tensors [P]."""
        if self._hist is None:
            raise RuntimeError("construct the suite with history=K to keep K readings per sensor")
        s = SENSOR_NAMES.index(name)
        ts = self._hist_times
        rows = []
        if ts:
            cutoff = ts[-1] - float(window_seconds)                       # base_sensor.py:771-775
            newest = self.read_index - 1
            rows = [(newest - j) % self.history for j in range(len(ts)) if ts[len(ts) - 1 - j] >= cutoff]
        dev = self.ens.device
        out = torch.empty((7, self.ens.n_plants), dtype=torch.float64, device=dev)
        rows_dev = torch.tensor(rows if rows else [0], dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.lib().wt_sensor_window_stats(self.ens.n_plants, len(rows), C.c_void_p(self._hist.data_ptr()),
                                                   C.c_void_p(rows_dev.data_ptr()), s, C.c_void_p(out.data_ptr()),
                                                   self._stream())
        _lib.check(rc, "wt_sensor_window_stats")
        return {k: out[i] for i, k in enumerate(("mean", "std", "min", "max", "count", "drift_rate", "fault_rate"))}

    def register_image(self, plants, sim_time: float):
        """Modbus input-register image of the selected plants from the LAST read:
        what update_modbus_inputs (__main__.py:166-224) hands to the Modbus slave, encoded on the device
        (modbus/protocols.py:34-58).  Returns (ir [K, 104] int16 holding the uint16 words, di [K, 3] uint8,
        ok [K] uint8); a row with ok == 0 is one the reference's update would have rejected."""
        dev = self.ens.device
        sel = torch.as_tensor(plants, dtype=torch.int32).reshape(-1).to(dev).contiguous()
        K = int(sel.numel())
        ir = torch.empty((K, 104), dtype=torch.int16, device=dev)
        di = torch.empty((K, 3), dtype=torch.uint8, device=dev)
        ok = torch.empty(K, dtype=torch.uint8, device=dev)
        p = lambda t: C.c_void_p(t.data_ptr())
        with torch.cuda.device(dev):
            rc = _lib.lib().wt_register_image(K, p(sel), self.ens.n_plants, p(self._out[0]), p(self._out_fault),
                                              float(sim_time), p(ir), p(di), p(ok), self._stream())
        _lib.check(rc, "wt_register_image")
        return ir, di, ok


def create_realistic_sensor_suite(ensemble, seed: int = 0, plant0: int = 0, history: int = 0, **suite_kw) -> SensorSuite:
    """Batched counterpart of sensors/__init__.py:41-120 for a PlantEnsemble (same 7 keys)."""
    return SensorSuite(ensemble, seed=seed, plant0=plant0, history=history, **suite_kw)
