"""Host-side pieces of the Python mirror that need no GPU."""
import numpy as np

from motion_planning_b200.mppi import _Log


def test_log_appends_like_concatenate():
    # the reference grows path / uvec with np.concatenate every step (control/src/mppi:95-97)
    rng = np.random.default_rng(0)
    rows = rng.normal(size=(300, 3))
    log = _Log(rows[0])
    ref = np.array([rows[0]])
    for r in rows[1:]:
        log.append(r)
        ref = np.concatenate((ref, np.array([r])))
    assert log.array.shape == (300, 3)
    assert np.array_equal(log.array, ref)
    assert np.array_equal(log.array[-1], rows[-1])


def test_log_from_array_and_row_copy():
    a = np.arange(8.0).reshape(4, 2)
    log = _Log.from_array(a)
    assert np.array_equal(log.array, a)
    row = np.array([1.0, 2.0])
    log.append(row)
    row[:] = 0.0                      # the log must hold a copy, not a reference to the caller's buffer
    assert np.array_equal(log.array[-1], [1.0, 2.0])
