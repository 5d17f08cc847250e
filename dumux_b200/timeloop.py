"""Host-side time-step control, mirroring dumux/common/timeloop.hh (same member meaning, python naming).

  TimeLoop             timeloop.hh:200-470   (setTimeStepSize :320-332, advanceTimeStep :239-252, finished :385-388,
                                              willBeFinished :394-397, maxTimeStepSize :404-411)
  CheckPointTimeLoop   timeloop.hh:480-700   (advanceTimeStep :527-563, maxTimeStepSize :570-576, periodic check points)
  run_instationary     the loop body of the reference mains (test/porousmediumflow/2p/incompressible/main.cc:133-163,
                       1p/compressible/instationary/main.cc:128-150) incl. NewtonSolver::solve's time-step halving
                       (nonlinear/newtonsolver.hh:309-355) and suggestTimeStepSize (:784-798)

The Newton step itself runs on the device (Engine) or in the CPU oracle; this module only decides dt.
"""
from __future__ import annotations


class TimeLoop:
    BASE_EPS = 1e-10

    def __init__(self, start_time: float, dt: float, t_end: float):
        self.start_time = start_time
        self.end_time = t_end
        self.time = start_time
        self.user_max_dt = float("inf")
        self.previous_dt = dt
        self.step_index = 0
        self.dt = 0.0
        self.set_time_step_size(dt)

    def finished(self) -> bool:
        return (self.end_time - self.time) < self.BASE_EPS * (self.time - self.start_time)

    def will_be_finished(self) -> bool:
        return self.finished() or (self.end_time - self.time - self.dt) < self.BASE_EPS * self.dt

    def max_time_step_size(self) -> float:
        if self.finished():
            return 0.0
        return min(self.user_max_dt, max(0.0, self.end_time - self.time))

    def set_time_step_size(self, dt: float):
        self.dt = min(dt, self.max_time_step_size())

    def set_max_time_step_size(self, max_dt: float):
        self.user_max_dt = max_dt
        self.set_time_step_size(self.dt)

    def advance_time_step(self):
        self.step_index += 1
        self.time += self.dt
        self.previous_dt = self.dt
        self.set_time_step_size(self.dt)


class CheckPointTimeLoop(TimeLoop):
    def __init__(self, start_time: float, dt: float, t_end: float):
        self.periodic = False
        self.delta_cp = 0.0
        self.last_cp = start_time
        self.is_check_point = False
        super().__init__(start_time, dt, t_end)

    # timeloop.hh:722-760: time to the next periodic check point
    def _dt_to_next_check_point(self, t: float) -> float:
        if not self.periodic:
            return float("inf")
        return self.last_cp + self.delta_cp - t

    def max_time_step_size(self) -> float:
        return min(super().max_time_step_size(), self._dt_to_next_check_point(self.time))

    def set_periodic_check_point(self, interval: float, offset: float = 0.0):
        self.periodic = True
        self.delta_cp = interval
        self.last_cp = offset
        # the first check point is the first one >= current time
        while self.last_cp + self.delta_cp < self.time + 1e-14 * self.delta_cp:
            self.last_cp += self.delta_cp
        self.set_time_step_size(self.dt)

    def advance_time_step(self):
        dt = self.dt
        new_time = self.time + dt
        # a periodic check point is hit when the new time reaches lastCheckPoint + delta (fuzzy equality as in FloatCmp::eq)
        hit = self.periodic and abs(new_time - (self.last_cp + self.delta_cp)) <= 1e-8 * self.delta_cp
        if hit:
            self.last_cp += self.delta_cp
        self.is_check_point = hit
        previous = self.previous_dt
        super().advance_time_step()
        if not self.will_be_finished():
            if self.is_check_point:
                self.set_time_step_size(max(dt, previous))
            next_dt = self.dt
            threshold = 0.2 * next_dt
            next_time = self.time + next_dt
            to_cp = self._dt_to_next_check_point(next_time)
            if 0.0 < to_cp <= threshold * (1 + 1e-8):
                next_dt += to_cp
            self.set_time_step_size(next_dt)


def suggest_time_step_size(dt: float, newton_iterations: int, target: int = 10) -> float:
    """NewtonSolver::suggestTimeStepSize, newtonsolver.hh:784-798"""
    if newton_iterations > target:
        return dt / (1.0 + (newton_iterations - target) / target)
    return dt * (1.0 + (target - newton_iterations) / target / 1.2)


def run_instationary(stepper, loop: TimeLoop, max_divisions: int = 10, target_steps: int = 10):
    """stepper: .solve(dt) -> (converged: bool, newton_iterations: int); .reset(); .advance().
    Returns (newton_iterations per step, dt per step)."""
    its, dts = [], []
    while True:
        n_it = 0
        ok = False
        for i in range(max_divisions + 1):
            ok, n_it = stepper.solve(loop.dt)
            if ok:
                break
            if i < max_divisions:
                stepper.reset()
                loop.set_time_step_size(loop.dt * 0.5)
        if not ok:
            raise RuntimeError(f"Newton solver didn't converge after {max_divisions} time-step divisions")
        stepper.advance()
        its.append(n_it)
        dts.append(loop.dt)
        loop.advance_time_step()
        loop.set_time_step_size(suggest_time_step_size(loop.dt, n_it, target_steps))
        if loop.finished():
            break
    return its, dts
