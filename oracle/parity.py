"""oracle/parity.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The comparison rules of the simulator parity checks, shared by tests/test_env_gpu.py,
tools/long_parity.py and __graft_entry__.smoke(): what "the kernel's step equals the oracle's"
means (``mismatch``), and the either-branch search (``either_branch``) that replaces skipping a
step on which the fp64 oracle took a decision within MARGIN of its threshold.
"""
from __future__ import annotations

import numpy as np

MARGIN = 1e-3
MARGIN64 = 1e-9   # fp64 kernel against the fp64 oracle: only exact ties (resting contacts, margins of 0 to 1e-13) may differ
TOL = dict(x=1e-3, y=1e-3, speed=2e-4, heading=1e-4, impact_x=1e-3, impact_y=1e-3, timer=1e-9, target_speed=1e-5)
# fp64 validation kernel against the fp64 oracle.  Positions / impacts 2e-6: when two nearly parallel rectangles
# collide, the separating-axis minimum between their (nearly identical) axes is a tie up to theta^2 ~ 1e-12 m that
# rounding decides; the impact (< 0.7 m) then points along one heading or the other, a difference of
# impact * theta <= ~1e-6 m.  Everything else is at rounding level (measured: 1e-12).
TOL64 = dict(x=2e-6, y=2e-6, speed=1e-7, heading=1e-7, impact_x=2e-6, impact_y=2e-6, timer=1e-9, target_speed=1e-9)
DISCRETE = ("lane", "target_lane", "crashed", "has_impact")
# Controlled vehicles acting below 0.5 m/s (jams behind a crash; IDM vehicles even reverse there): the steering
# controller divides by the speed twice and a reversing vehicle's lateral loop is unstable, so a 1e-7 rounding of y
# grows by orders of magnitude within one policy step (measured: a reversing vehicle's heading off by 0.3 rad, its
# follower's speed by 1.4e-2 m/s).  Such a vehicle's continuous tolerances are widened by SLOW_FACTOR, and so are
# those of the vehicles within NEAR_SLOW_M metres of it (they follow it: its position enters their IDM gap).
# Nothing discrete is relaxed.
SLOW_FACTOR = 1e3
NEAR_SLOW_FACTOR = 1e3
NEAR_SLOW_M = 60.0
# A vehicle that REVERSES is laterally unstable under highway-env's own steering law: its deviation from the lane centre
# and its heading grow by a factor 5 to 8 per simulation frame, in the oracle exactly as in the kernel, and a rounding
# difference between the two grows at the same rate (profiles/r02_reversing_vehicle_trace.txt: fp64 kernel against fp64
# oracle, 2e-16 m after frame 0, 6e-4 m after frame 14, while the oracle's own y goes from 2e-8 m to 5e-2 m).  Its error
# is therefore bounded RELATIVE to the deviation the oracle itself reports at the end of the step, not absolutely.
SLOW_REL = 0.05


class Got:
    """One env's slice of a kernel step's results."""

    def __init__(self, state, e, obs, rew, term, trunc, rows):
        self.state = {k: state[k][e] for k in list(TOL) + list(DISCRETE)}
        self.obs, self.rew, self.term, self.trunc, self.rows = obs[e], float(rew[e]), bool(term[e]), bool(trunc[e]), rows[e]


def compare(got, o, r, te, tr, want_obs, want_rows, tol, obs_tol, rew_tol, worst=None):
    """None when the kernel's results equal this oracle outcome in full, else a dict: `why` (what differs first),
    `vehicles` (whose state differs) and `score` (how much differs: the search descends on it)."""
    ref = o.get_state()
    bad, why = set(), None
    for k in DISCRETE:
        d = np.nonzero(got.state[k] != ref[k])[0]
        if len(d):
            bad.update(int(v) for v in d)
            why = why or f"{k}: vehicles {d.tolist()}"
    score = 10 * len(bad)
    if got.term != te or got.trunc != tr:
        score += 10
        why = why or "terminated/truncated"
    if not np.array_equal(got.rows, want_rows):
        score += 5
        why = why or "observation rows"
    slow = o.slow_vehicles()
    scale = np.ones(len(slow))
    if slow.any():
        near = np.abs(ref["x"][:, None] - ref["x"][None, slow]).min(axis=1) < NEAR_SLOW_M
        scale = np.where(slow, SLOW_FACTOR, np.where(near, NEAR_SLOW_FACTOR, 1.0))
    errs = {}
    rel = {}
    if slow.any():
        rel = {"y": SLOW_REL * np.abs(ref["y"] - 4.0 * ref["lane"]) * slow, "heading": SLOW_REL * np.abs(ref["heading"]) * slow}
    for k, t in tol.items():
        err = np.maximum(np.abs(got.state[k] - ref[k]) - rel.get(k, 0.0), 0.0) / scale
        errs[k] = float(err.max())
        over = np.nonzero(err > t)[0]
        if len(over):
            score += len(over)
            bad.update(int(v) for v in over)
            why = why or f"{k} {errs[k]:.3g} > {t:g}: vehicle {int(err.argmax())}"
    errs["reward"] = abs(got.rew - r)
    if errs["reward"] > rew_tol:
        score += 1
        why = why or f"reward {got.rew} vs {r}"
    errs["obs"] = float(np.max(np.abs(got.obs - want_obs)))
    if errs["obs"] > obs_tol * (SLOW_FACTOR if slow.any() else 1.0):
        score += 1
        why = why or f"observation {errs['obs']:.3g}"
    if why is None:
        if worst is not None:
            for k, v in errs.items():
                worst[k] = max(worst.get(k, 0.0), v)
        return None
    return {"why": why, "vehicles": bad, "score": score}


def mismatch(got, o, r, te, tr, want_obs, want_rows, tol, obs_tol, rew_tol, worst=None):
    """None when the kernel's results equal this oracle outcome in full, else what differs first."""
    c = compare(got, o, r, te, tr, want_obs, want_rows, tol, obs_tol, rew_tol, worst)
    return None if c is None else c["why"]


# kinds whose a / b fields are vehicle indices (x_order, abort_*, precheck, sat_*: both; the rest: a only)
_PAIR_KINDS = {5, 9, 10, 14, 15, 16, 17, 20}


def _priority(key, vehicles):
    """Search order of a marginal decision: decisions about the vehicles whose state differs first, then later
    frames first (the discrete state at the end of a step is decided by its last frames)."""
    from . import highway as oh

    kind, frame, a, b, _ = oh.key_fields(key)
    who = {a, b} if kind in _PAIR_KINDS else {a}
    return (0 if (vehicles and who & vehicles) else 1, -frame if frame < 250 else -999)


def either_branch(o, st0, action, perm, got, tol, obs_tol, rew_tol, max_nodes=3000, max_depth=12, why=None, margin=MARGIN):
    """Re-run the oracle step from ``st0`` with marginal decisions forced the other way; returns the forced keys
    under which the kernel's results equal the oracle's in full, or None.  Best-first search over sets of forced
    decisions: the node whose outcome differs least from the kernel's is expanded first (a flip that repairs one of
    several differing vehicles leads on), its children ordered by `_priority`.  A flip that leaves the oracle's
    outcome bit-identical to its parent's (e.g. the contact test of two vehicles that have crashed already) is
    not extended, one that takes the outcome further from the kernel's is not extended either."""
    import heapq

    o.record_margin(margin)
    seen, heap, found, nodes, tick = {()}, [], None, 0, 0

    def run(forced):
        o.set_state(st0)
        o.force(forced)
        r, te, tr = o.step(action)
        want_obs, want_rows = o.observe(perm=perm, with_rows=True)
        st = o.get_state()
        sig = hash(tuple(st[k].tobytes() for k in ("x", "y", "speed", "heading") + DISCRETE) + (want_rows.tobytes(),))
        return compare(got, o, r, te, tr, want_obs, want_rows, tol, obs_tol, rew_tol), o.marginal(), sig

    c, keys, sig = run(())
    if c is not None:
        heapq.heappush(heap, (c["score"], 0, tick, (), c, keys, sig))
    while heap and nodes < max_nodes and found is None:
        _, depth, _, forced, c, keys, sig = heapq.heappop(heap)
        if depth >= max_depth:
            continue
        for k in sorted((k for k in keys if k not in forced), key=lambda k: _priority(k, c["vehicles"])):
            nxt = tuple(sorted(forced + (k,)))
            if nxt in seen:
                continue
            seen.add(nxt)
            nodes += 1
            c2, keys2, sig2 = run(nxt)
            if c2 is None:
                found = nxt
                break
            if sig2 != sig and c2["score"] <= c["score"]:
                tick += 1
                heapq.heappush(heap, (c2["score"], depth + 1, tick, nxt, c2, keys2, sig2))
            if nodes >= max_nodes:
                break
    # leave the oracle on its own (unforced) outcome
    o.force(())
    o.record_margin(0.0)
    o.set_state(st0)
    o.step(action)
    return found


# ---- frame-by-frame either-branch search ------------------------------------------------------------------------
# A pile-up of crashed vehicles is a chain of resting contacts: after an impact translation two rectangles touch
# EXACTLY, so the contact tests of the following frames are decided by the last bit (margins of 0 to 1e-13 in the
# oracle), every frame anew, and every outcome changes the states the next frame's tests see.  The set of decisions
# that reproduces the kernel's trajectory cannot be found from the step's final state alone.  With the kernel's
# per-frame trace (hrp_env_set_trace) and the oracle's (hw_set_trace) it can: walk the frames, and at the first frame
# whose snapshots differ, flip marginal decisions OF THAT FRAME until the snapshots agree, then go on.
_TRACE_TOL = np.array([TOL["x"], TOL["y"], TOL["speed"], TOL["heading"], TOL["impact_x"], TOL["impact_y"]])


def _first_diff(ktrace, otrace, scale, ttol=None):
    """(frame, vehicles) of the first per-frame snapshot that differs between kernel and oracle, or None."""
    ttol = _TRACE_TOL if ttol is None else ttol
    for f in range(otrace.shape[0]):
        bad = np.nonzero(ktrace[f, :, 6] != otrace[f, :, 6])[0]
        err = np.abs(ktrace[f, :, :6] - otrace[f, :, :6]) / scale[:, None]
        over = np.nonzero((err > ttol[None, :]).any(axis=1))[0]
        if len(bad) or len(over):
            return f, {int(v) for v in bad} | {int(v) for v in over}
    return None


def frame_search(o, st0, action, perm, got, ktrace, tol, obs_tol, rew_tol, max_runs=1500, margin=MARGIN, trace_tol=None):
    """Either-branch search guided by the kernel's per-frame trace ``ktrace`` [frames, V, 7].  Returns the forced keys
    under which the oracle's step equals the kernel's in full (every frame's snapshot and the final results), or None."""
    from . import highway as oh

    otrace = o.trace(True)
    o.record_margin(margin)
    runs = 0
    ttol = _TRACE_TOL if trace_tol is None else trace_tol

    def run(forced):
        nonlocal runs
        runs += 1
        o.set_state(st0)
        o.force(forced)
        r, te, tr = o.step(action)
        want_obs, want_rows = o.observe(perm=perm, with_rows=True)
        slow = o.slow_vehicles()
        scale = np.ones(len(slow))
        if slow.any():
            x = o.get_state()["x"]
            near = np.abs(x[:, None] - x[None, slow]).min(axis=1) < NEAR_SLOW_M
            scale = np.where(slow | near, SLOW_FACTOR, 1.0)
        return (_first_diff(ktrace, otrace, scale, ttol), compare(got, o, r, te, tr, want_obs, want_rows, tol, obs_tol, rew_tol),
                o.marginal())

    forced, found = (), None
    diff, c, keys = run(forced)
    while runs < max_runs:
        if diff is None:
            if c is None:
                found = forced
                break
            f = 250   # every frame agrees: what is left are the decisions after the last frame (on-road test, observation)
        else:
            f = diff[0]
        in_frame = (lambda k: oh.key_fields(k)[1] >= 250) if f == 250 else (lambda k: oh.key_fields(k)[1] == f)
        # breadth-first over the decisions of frame f: a forced decision can expose further ones of the same frame (the
        # separating-axis loop stops at the first separating axis; once that is flipped the next axis is evaluated)
        queue, seen, step = [(forced, diff, c, keys)], {forced}, None
        while queue and step is None and runs < max_runs:
            base, bdiff, bc, bkeys = queue.pop(0)
            vehicles = bdiff[1] if bdiff is not None else (bc["vehicles"] if bc else set())
            for k in sorted((k for k in bkeys if in_frame(k) and k not in base), key=lambda k: _priority(k, vehicles)):
                nxt = tuple(sorted(base + (k,)))
                if nxt in seen:
                    continue
                seen.add(nxt)
                d2, c2, k2 = run(nxt)
                if (d2 is None and (f != 250 or c2 is None)) or (d2 is not None and f != 250 and d2[0] > f):
                    step = (nxt, d2, c2, k2)
                    break
                if len(nxt) - len(forced) < 6 and (f == 250 or (d2 is not None and d2[0] == f)):
                    queue.append((nxt, d2, c2, k2))
                if runs >= max_runs:
                    break
        if step is None:
            break
        forced, diff, c, keys = step
    o.trace(False)
    o.force(())
    o.record_margin(0.0)
    o.set_state(st0)
    o.step(action)
    return found
