"""GPU: behaviour-level checks of the whole controller path (properties that hold at any size)."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
PARK = np.array([0.0, -1.0, 0.0])


def mp():
    import motion_planning_b200 as m
    return m


def test_solve_path_reaches_the_goal_like_the_reference_loop():
    """MPPI.solve_path (control/src/mppi:104-125): closed loop on the model until within `thresh`."""
    m = mp().MPPI(horizon=32, samples=4096, seed=1)
    start, goal = np.array([0.0, 0.0, np.pi / 2.0]), np.array([0.3, 0.4, np.pi / 2.0])
    m.solve_path(start, goal, max_iters=2000)
    assert np.linalg.norm(m.path[-1][:2] - goal[:2]) <= m.thresh
    assert np.all(np.abs(m.uvec) <= 6.35492 + 1e-12)
    assert m.path.shape[0] == m.uvec.shape[0] == len(m.fin_time)
    m.close()


def test_steps_are_deterministic_and_restartable():
    """Same seed -> bit-identical runs; get/set of latest_uvec + use_philox(seed) restarts a run exactly."""
    K, T = 8192, 32

    def run(n, m=None):
        m = m or mp().MPPI(horizon=T, samples=K, seed=5)
        s = np.array([0.1, 0.0, 0.3])
        outs = []
        for _ in range(n):
            s = m.get_path(s, PARK)
            outs.append((m.uvec[-1].copy(), s.copy(), m.latest_uvec))
        return m, outs

    a, oa = run(4)
    b, ob = run(4)
    for (u1, s1, U1), (u2, s2, U2) in zip(oa, ob):
        assert np.array_equal(u1, u2) and np.array_equal(s1, s2) and np.array_equal(U1, U2)
    a.close()
    b.close()


def test_full_size_properties_c2():
    """BASELINE config 2 size: controls stay inside the clip, the last column is the shifted-in zero,
    the applied control equals column 0 of the pre-shift sequence, noise moments are right."""
    K, T = 65536, 64
    m = mp().MPPI(horizon=T, samples=K, seed=0)
    s = np.zeros(3)
    for it in range(3):
        s = m.get_path(s, PARK)
        U, Upre = m.latest_uvec, m.get_last_update()
        assert np.all(np.abs(U) <= 6.35492) and np.all(np.isfinite(U))
        assert np.all(U[:, -1] == 0.0) and np.array_equal(U[:, :-1], Upre[:, 1:])      # control/src/mppi:100-101
        assert np.array_equal(m.uvec[-1], Upre[:, 0])                                   # control/src/mppi:96-97
    eps = m.get_noise()
    assert eps.shape == (T, 2, K)
    assert abs(eps.mean()) < 1e-3 and abs(eps.std() - 0.9) < 1e-3
    st = m.stats()
    assert st["refine_overflow"] == 0 and 0 < st["refine_candidates"] < 64 * 32
    assert st["refine_max_dev"] < 0.01       # fp32 screen vs fp64 re-evaluation, head-room is 0.04
    m.close()


def test_cpp_facade_closed_loop():
    """include/mppi.hpp: compile examples/ros_control_node.cpp against the built library and run it."""
    exe = "/tmp/mppi_ros_node_test"
    lib = os.path.join(ROOT, "motion_planning_b200", "lib")
    r = subprocess.run(["g++", "-std=c++17", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "ros_control_node.cpp"),
                        "-L" + lib, "-lmppi_b200", "-Wl,-rpath," + lib, "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, "2048", "32", "1500"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "WAYPOINT REACHED" in r.stdout
    # the same node with the robot handed over as a user ODE functor (mppi::UserDynamics -> mppi_create_user, NVRTC)
    r = subprocess.run([exe, "2048", "32", "600", "user"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "user-supplied ODE functor" in r.stdout and "WAYPOINT REACHED" in r.stdout
    # ... and as a kinematic functor (mppi::UserKinematics, precision MIXED)
    r = subprocess.run([exe, "2048", "32", "600", "kinematic"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "kinematic functor" in r.stdout and "WAYPOINT REACHED" in r.stdout


def test_mixed_overflow_backs_off_to_fp64_and_stays_exact():
    """A candidate-list overflow of the mixed mode (here forced by an absurd screening margin; in the field: within
    centimetres of the goal, where hundreds of rollouts carry weight) redoes the step in fp64 and sends the next 8, 16,
    32 ... steps straight to the fp64 pipeline.  The controls must equal those of an fp64 engine, and the number of
    doomed mixed attempts must follow the back-off schedule (steps 1, 10, 27 of 40)."""
    K, T = 4096, 32
    a = mp().MPPI(horizon=T, samples=K, seed=6, precision="mixed", refine_margin=50.0)
    b = mp().MPPI(horizon=T, samples=K, seed=6, precision="f64")
    sa = sb = np.array([0.0, 0.0, 0.2])
    for _ in range(40):
        sa = a.get_path(sa, PARK)
        sb = b.get_path(sb, PARK)
        np.testing.assert_allclose(a.latest_uvec, b.latest_uvec, rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(sa, sb, rtol=1e-12, atol=1e-14)
    assert a.stats()["refine_overflow"] == 3
    a.initialize()                      # initialize() re-arms the short hold-off
    a.get_path(np.zeros(3), PARK)
    a.get_path(np.zeros(3), PARK)
    assert a.stats()["refine_overflow"] == 4
    a.close()
    b.close()


def test_sharded_external_exchange_survives_the_overflow_regime():
    """a sharded engine with an EXTERNAL exchange (mppi_step_local / exchange / mppi_step_finish, the 'nccl' and
    'host' transports of ShardedMPPI) must not die when the fp32 screen's candidate lists overflow (systematic within
    centimetres of the goal for large K): the split-phase API answers MPPI_ERR_RETRY on EVERY rank, nothing is applied, and the
    repeated round trip runs in fp64 on the same noise, followed by the same hold-off schedule as mppi_step.  The overflow is
    provoked deterministically here by a screening window of 50 cost units (thousands of rollouts inside it); two shards on one
    device, host-staged exchange; a single fp64 engine is the reference trajectory."""
    import motion_planning_b200 as mp
    from motion_planning_b200 import _capi
    from motion_planning_b200.distributed import shard_plan
    K, T, G = 8192, 32, 2
    goal = np.array([0.0, -1.0, 0.0])
    one = mp.MPPI(horizon=T, samples=K, precision="f64", seed=0)
    eng = []
    for r in range(G):
        kl, ko = shard_plan(K, G, r)
        eng.append(mp.MPPI(horizon=T, samples=kl, precision="mixed", seed=0, k_offset=ko, k_total=K, world_size=G, rank=r,
                           refine_margin=50.0))
    s, retries, worst, attempts_log = np.zeros(3), 0, 0.0, []
    for it in range(24):
        s1 = one.get_path(s, goal)
        for attempt in range(3):
            recs = []
            for e in eng:
                _capi.check(e._lib.mppi_set_goal(e._h, _capi.dptr(goal)), "goal")
                _capi.check(e._lib.mppi_step_local(e._h, _capi.dptr(_capi.f64(s))), "local")
                rec = np.empty(T * 6)
                _capi.check(e._lib.mppi_read_record(e._h, _capi.dptr(rec)), "read")
                recs.append(rec)
            allrec = np.concatenate(recs)
            sts = []
            for e in eng:
                _capi.check(e._lib.mppi_write_gather(e._h, _capi.dptr(allrec)), "write")
                u, x = np.empty(2), np.empty(3)
                sts.append(e._lib.mppi_step_finish(e._h, _capi.dptr(u), _capi.dptr(x)))
            assert len(set(sts)) == 1, sts                       # every rank takes the same decision
            if sts[0] != _capi.MPPI_ERR_RETRY:
                break
            retries += 1
        attempts_log.append(attempt)
        _capi.check(sts[0], "mppi_step_finish")
        errU = max(np.max(np.abs(e.latest_uvec - one.latest_uvec)) for e in eng) / max(1.0, np.max(np.abs(one.latest_uvec)))
        worst = max(worst, errU)
        assert errU < 1e-8, (it, errU)
        np.testing.assert_allclose(x, s1, rtol=0, atol=1e-10)
        assert np.array_equal(eng[0].latest_uvec, eng[1].latest_uvec)      # every rank holds the identical nominal
        # every step is compared from IDENTICAL inputs: the map U -> U' amplifies a 1e-10 difference of the nominal by up to
        # 1/lam per step (the soft-min is nearly an arg-min), so two free-running loops drift apart chaotically
        for e in eng:
            e.latest_uvec = one.latest_uvec
        s = s1
    print("sharded external exchange: attempts per step %s, worst rel err U %.3g" % (attempts_log, worst))
    # step 0 overflows and is retried in fp64; the next 8 steps run fp64 directly (hold-off), step 9 probes mixed again, ...
    assert attempts_log[0] == 1 and attempts_log[1:9] == [0] * 8 and attempts_log[9] == 1 and retries >= 2
    for o in [one] + eng:
        o.close()
