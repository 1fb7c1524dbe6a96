"""ctypes binding of include/rm_radar_b200.h.  Fails loudly when the CUDA library is missing —
there is no CPU fallback anywhere in the product path."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# RMR_LIB_PATH: another build of the same library (A/B timing of two kernel versions on one GPU box); default in-tree
LIB_PATH = os.environ.get("RMR_LIB_PATH") or os.path.join(HERE, "librm_radar_b200.so")
MAX_ARMORS = 16


class Detection(C.Structure):
    # radar::Detection — /root/reference/src/detect/detection.h:25-68
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("width", C.c_float), ("height", C.c_float),
                ("label", C.c_float), ("confidence", C.c_float)]

    def astuple(self):
        return (self.x, self.y, self.width, self.height, self.label, self.confidence)


class RobotRec(C.Structure):
    _fields_ = [("rect", C.c_float * 4), ("has_rect", C.c_int32), ("is_detected", C.c_int32),
                ("label", C.c_int32), ("confidence", C.c_float), ("n_armors", C.c_int32),
                ("armors", Detection * MAX_ARMORS), ("is_located", C.c_int32), ("location", C.c_float * 3),
                ("cluster", C.c_int32), ("cluster_points", C.c_int32)]


class TrackRec(C.Structure):
    _fields_ = [("id", C.c_int32), ("label", C.c_int32), ("state", C.c_int32), ("init_count", C.c_int32),
                ("miss_count", C.c_int32), ("location", C.c_float * 3), ("filter_state", C.c_float * 9)]


# every symbol include/rm_radar_b200.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "rmr_last_error", "rmr_device_count",
    "rmr_detector_create", "rmr_detector_destroy", "rmr_detector_detect", "rmr_detector_detect_batch",
    "rmr_detector_last_input", "rmr_detector_last_output", "rmr_detector_info", "rmr_detector_set_stream",
    "rmr_detector_time_forward", "rmr_detector_profile_ops", "rmr_detector_plan_stats",
    "rmr_robot_detector_create", "rmr_robot_detector_create_batched", "rmr_robot_detector_detect_frames", "rmr_run_batch", "rmr_robot_detector_destroy", "rmr_robot_detector_detect",
    "rmr_robot_detector_detect_device", "rmr_robot_detector_last_cars", "rmr_robot_detector_last_armors",
    "rmr_robot_detector_set_stream", "rmr_robot_detector_last_stats", "rmr_robot_detector_last_timing", "rmr_robot_detector_car",
    "rmr_robot_detector_armor",
    "rmr_locator_create", "rmr_locator_destroy", "rmr_locator_update", "rmr_locator_update_device",
    "rmr_locator_cluster", "rmr_locator_search", "rmr_locator_update_pcd", "rmr_pcd_parse", "rmr_locator_load_background", "rmr_locator_set_stream", "rmr_locator_image_size",
    "rmr_locator_read_image", "rmr_locator_stats", "rmr_locator_read_foreground",
    "rmr_run_once", "rmr_conv_selftest", "rmr_conv_timeline", "rmr_conv_plan", "rmr_postprocess_selftest", "rmr_engine_build", "rmr_engine_resolve",
    "rmr_comm_unique_id", "rmr_comm_create", "rmr_comm_close", "rmr_comm_destroy", "rmr_comm_publish", "rmr_comm_collect", "rmr_comm_pack",
    "rmr_tracker_create", "rmr_tracker_destroy", "rmr_tracker_update", "rmr_tracker_tracks", "rmr_auction",
    "rmr_jpeg_decoder_create", "rmr_jpeg_decoder_destroy", "rmr_jpeg_decoder_set_stream", "rmr_jpeg_info", "rmr_jpeg_decode",
    "rmr_jpeg_decode_device", "rmr_jpeg_decoder_status", "rmr_jpeg_decoder_read_coefficients", "rmr_jpeg_decoder_profile", "rmr_robot_detector_detect_jpeg",
]

_lib = None


class RadarError(RuntimeError):
    pass


class CapacityError(RadarError):
    """RMR_ERR_CAPACITY: a fixed internal capacity was exceeded; the call fails instead of returning a truncated result."""


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RadarError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(rm_radar_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.rmr_last_error.restype = C.c_char_p
    vp, ci, cf, cd = C.c_void_p, C.c_int, C.c_float, C.c_double
    P = C.POINTER
    lib.rmr_device_count.argtypes = [P(ci)]
    lib.rmr_detector_create.argtypes = [P(vp), C.c_char_p, ci, ci, ci, ci, cf, cf, ci, ci, ci, ci]
    lib.rmr_detector_destroy.argtypes = [vp]
    lib.rmr_detector_destroy.restype = None
    lib.rmr_detector_detect.argtypes = [vp, vp, ci, ci, ci, P(Detection), ci, P(ci)]
    lib.rmr_detector_detect_batch.argtypes = [vp, P(vp), P(ci), P(ci), P(ci), ci, P(Detection), ci, P(ci)]
    lib.rmr_detector_last_input.argtypes = [vp, vp, ci]
    lib.rmr_detector_last_output.argtypes = [vp, vp, ci]
    lib.rmr_detector_info.argtypes = [vp, P(ci), P(ci), P(ci), P(cd)]
    lib.rmr_detector_set_stream.argtypes = [vp, vp]
    lib.rmr_detector_time_forward.argtypes = [vp, ci, ci, P(cf)]
    lib.rmr_detector_plan_stats.argtypes = [vp, ci, P(ci), P(ci), P(ci)]
    lib.rmr_detector_profile_ops.argtypes = [vp, ci, ci, P(cd), ci, P(ci)]
    lib.rmr_robot_detector_create.argtypes = [P(vp), C.c_char_p, C.c_char_p, ci, ci, ci, ci, cf, cf, cf, cf, cf,
                                              ci, ci, ci, ci]
    lib.rmr_robot_detector_create_batched.argtypes = [P(vp), C.c_char_p, C.c_char_p, ci, ci, ci, ci, cf, cf, cf, cf, cf,
                                                      ci, ci, ci, ci, ci]
    lib.rmr_robot_detector_detect_frames.argtypes = [vp, vp, ci, ci, ci, ci, ci, vp, ci, P(ci)]
    lib.rmr_run_batch.argtypes = [vp, P(vp), ci, vp, ci, ci, ci, ci, vp, ci, ci, ci, vp, ci, P(ci)]
    lib.rmr_robot_detector_destroy.argtypes = [vp]
    lib.rmr_robot_detector_destroy.restype = None
    lib.rmr_robot_detector_detect.argtypes = [vp, vp, ci, ci, ci, P(RobotRec), ci, P(ci)]
    lib.rmr_robot_detector_detect_device.argtypes = [vp, vp, ci, ci, ci, P(RobotRec), ci, P(ci)]
    lib.rmr_robot_detector_last_cars.argtypes = [vp, P(Detection), ci, P(ci)]
    lib.rmr_robot_detector_last_armors.argtypes = [vp, ci, P(Detection), ci, P(ci)]
    lib.rmr_robot_detector_set_stream.argtypes = [vp, vp]
    lib.rmr_robot_detector_last_stats.argtypes = [vp, P(ci), P(cd), P(ci)]
    lib.rmr_robot_detector_last_timing.argtypes = [vp, P(cf), P(cf)]
    lib.rmr_robot_detector_car.argtypes = [vp]
    lib.rmr_robot_detector_car.restype = vp
    lib.rmr_robot_detector_armor.argtypes = [vp]
    lib.rmr_robot_detector_armor.restype = vp
    lib.rmr_locator_create.argtypes = [P(vp), ci, ci, P(cf), P(cf), P(cf), cf, ci, cf, cf, cf, ci, ci, cf, ci]
    lib.rmr_locator_destroy.argtypes = [vp]
    lib.rmr_locator_destroy.restype = None
    lib.rmr_locator_update.argtypes = [vp, vp, ci, ci]
    lib.rmr_locator_update_device.argtypes = [vp, vp, ci, ci]
    lib.rmr_locator_cluster.argtypes = [vp]
    lib.rmr_locator_search.argtypes = [vp, P(RobotRec), ci]
    lib.rmr_locator_update_pcd.argtypes = [vp, vp, C.c_size_t, P(ci)]
    lib.rmr_pcd_parse.argtypes = [vp, C.c_size_t, vp, ci, P(ci), ci]
    lib.rmr_tracker_create.argtypes = [P(vp), P(cf), ci, ci, ci, cf, cf, cf, cf, ci, cf]
    lib.rmr_tracker_destroy.argtypes = [vp]
    lib.rmr_tracker_destroy.restype = None
    lib.rmr_tracker_update.argtypes = [vp, vp, ci, C.c_int64, P(C.c_int32), P(C.c_int32)]
    lib.rmr_tracker_tracks.argtypes = [vp, vp, ci, P(ci)]
    lib.rmr_auction.argtypes = [vp, ci, ci, ci, P(C.c_int32)]
    lib.rmr_jpeg_decoder_create.argtypes = [P(vp), ci]
    lib.rmr_jpeg_decoder_destroy.argtypes = [vp]
    lib.rmr_jpeg_decoder_destroy.restype = None
    lib.rmr_jpeg_decoder_set_stream.argtypes = [vp, vp]
    lib.rmr_jpeg_info.argtypes = [vp, C.c_size_t] + [P(ci)] * 6
    lib.rmr_jpeg_decode.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, P(ci), P(ci)]
    lib.rmr_jpeg_decode_device.argtypes = [vp, vp, C.c_size_t, vp, ci, P(vp), P(ci), P(ci)]
    lib.rmr_jpeg_decoder_status.argtypes = [vp, P(ci), P(ci), P(ci), P(ci), P(C.c_size_t)]
    lib.rmr_jpeg_decoder_profile.argtypes = [vp, vp, C.c_size_t, P(cf)]
    lib.rmr_jpeg_decoder_read_coefficients.argtypes = [vp, vp, C.c_long, P(C.c_long)]
    lib.rmr_robot_detector_detect_jpeg.argtypes = [vp, vp, vp, C.c_size_t, vp, ci, P(ci)]
    lib.rmr_locator_load_background.argtypes = [vp, vp, ci, ci]
    lib.rmr_locator_set_stream.argtypes = [vp, vp]
    lib.rmr_locator_image_size.argtypes = [vp, P(ci), P(ci)]
    lib.rmr_locator_read_image.argtypes = [vp, ci, vp]
    lib.rmr_locator_stats.argtypes = [vp, P(ci), P(ci)]
    lib.rmr_locator_read_foreground.argtypes = [vp, vp, ci]
    lib.rmr_conv_selftest.argtypes = [ci, ci, ci, ci, ci, ci, ci, ci, ci, ci, C.c_uint, ci, P(cf), P(cf), P(cf)]
    lib.rmr_run_once.argtypes = [vp, vp, vp, ci, ci, ci, ci, vp, ci, ci, ci, P(RobotRec), ci, P(ci)]
    lib.rmr_conv_timeline.argtypes = [ci, ci, ci, ci, ci, ci, ci, vp, ci, P(ci)]
    lib.rmr_conv_plan.argtypes = [ci, ci, ci, ci, ci, ci, ci, P(ci)]
    lib.rmr_engine_build.argtypes = [C.c_char_p, C.c_char_p, ci, ci]
    lib.rmr_engine_resolve.argtypes = [C.c_char_p, ci, ci, C.c_char_p, ci]
    lib.rmr_postprocess_selftest.argtypes = [vp, ci, cf, vp, ci, P(ci)]
    lib.rmr_comm_unique_id.argtypes = [vp]
    lib.rmr_comm_create.argtypes = [P(vp), vp, ci, ci, ci, ci]
    lib.rmr_comm_close.argtypes = [vp]
    lib.rmr_comm_destroy.argtypes = [vp]
    lib.rmr_comm_destroy.restype = None
    lib.rmr_comm_publish.argtypes = [vp, vp, ci, vp]
    lib.rmr_comm_collect.argtypes = [vp, vp]
    lib.rmr_comm_pack.argtypes = [vp, ci, ci, vp]
    _lib = lib
    return lib


def check(status: int):
    """Constructors throw in the reference (std::invalid_argument / std::runtime_error); hot-path
    CUDA failures abort.  Here every failure raises."""
    if status == 0:
        return
    msg = load().rmr_last_error().decode(errors="replace")
    if status == -1:
        raise ValueError(msg)
    if status == -4:
        raise CapacityError(msg)
    raise RadarError(f"rm_radar_b200 status {status}: {msg}")
