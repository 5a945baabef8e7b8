"""Sliding-window driver over the B200 bundle adjuster (mirrors pysfm's window_slam.py:17-67).

For every window of `window_size` consecutive cameras: set_bundle on that camera subset and the
first `num_tracks` tracks, optimize() on the device, then propagate the pose update of the
window's first camera exactly as the reference does.  The reference's plotting
(draw_bundle_pca / matplotlib, window_slam.py:49-67) is not part of the path: `pdf_pattern` is
accepted for signature compatibility and ignored.

Reference quirks kept on purpose (parity is checked against the unmodified reference):
  * `camera_mask = arange(window_size) < num_to_freeze` is computed but never handed to
    set_bundle (window_slam.py:24,37-38), so only the first camera of each window is fixed;
  * the propagation call passes the window's first camera as the camera to update
    (window_slam.py:44-46).
"""
from copy import deepcopy

import numpy as np

from . import geometry
from .bundle_adjuster import BundleAdjuster

NUM_TRACKS = 100   # window_slam.py:18


def run(complete_bundle, window_size, num_to_freeze=2, pdf_pattern=None, num_tracks=NUM_TRACKS,
        device=None, verbose=True, on_window=None):
    """Returns the adjusted bundle.  `on_window(i, adjuster)` is called after each window (the hook
    the reference uses for drawing)."""
    camera_mask = np.arange(window_size) < num_to_freeze   # noqa: F841  (unused in the reference too)
    track_ids = list(range(num_tracks))
    cur_bundle = complete_bundle
    # ONE adjuster for the whole run.  The measurements of the complete bundle are uploaded by the
    # first set_bundle and stay on the device (every later bundle shares them: optimize() clones
    # parameters, not tracks); a window costs two look-up tables, the device packer and the window's
    # camera / point parameters.  The reference deep-copies the bundle per window only to keep the
    # pose of the window's first camera (window_slam.py:32,44): that camera is copied instead.
    ba = BundleAdjuster(device=device, verbose=verbose)
    for i in range(0, len(complete_bundle.cameras) - window_size + 1):
        if verbose:
            print('\n\n==============\nWINDOW: [%d..%d]\n' % (i, i + window_size))
        prev_camera = deepcopy(cur_bundle.cameras[i])
        camera_ids = list(range(i, i + window_size))
        ba.set_bundle(cur_bundle, camera_ids=camera_ids, track_ids=track_ids)
        ba.optimize()
        cur_bundle = ba.bundle
        next_camera_id = i + window_size
        if next_camera_id < len(cur_bundle.cameras):
            geometry.propagate_pose_update_inplace(prev_camera, cur_bundle.cameras[i],
                                                   cur_bundle.cameras[i])
        if on_window is not None:
            on_window(i, ba)
    return cur_bundle


if __name__ == '__main__':
    import sys
    from . import bundle_io
    window_size = int(sys.argv[3])
    print('Loading bundle...')
    bundle = bundle_io.load(sys.argv[1], sys.argv[2])
    print('Triangulating initial points...')
    bundle.triangulate_all()          # one kernel launch (ba_triangulate)
    print('Cameras:', len(bundle.cameras))
    print('Tracks:', len(bundle.tracks))
    print('Window Size:', window_size)
    run(bundle, window_size)
