"""`SolverBase`: constructor contract, event preprocessing and the abstract `estimate`
(src/solver/base.py:54-150).  Visualisation helpers of the upstream class are outside the
contrast-maximisation path and are not reproduced."""
import logging
from typing import List, Optional, Tuple

import numpy as np
import torch

from .. import event_image_converter, utils, warp

logger = logging.getLogger(__name__)

# optimiser names the upstream solvers accept (src/solver/base.py:20-51)
SCIPY_OPTIMIZERS = ["Nelder-Mead", "Powell", "CG", "BFGS", "Newton-CG", "L-BFGS-B", "TNC", "COBYLA", "SLSQP",
                    "trust-constr", "dogleg", "trust-ncg", "trust-exact", "trust-krylov"]
TORCH_OPTIMIZERS = ["Adadelta", "Adagrad", "Adam", "AdamW", "SparseAdam", "Adamax", "ASGD", "LBFGS", "NAdam", "RAdam",
                    "RMSprop", "Rprop", "SGD"]


class SolverBase(object):
    """Base class for solvers.

    Params:
        orig_image_shape (tuple) ... (H, W) of the sensor.
        crop_image_shape (tuple) ... (H, W) of the region of interest.
        calibration_parameter (dict) ... stored only.
        solver_config (dict) ... the `solver:` block of the yaml config.
        visualize_module ... kept for signature parity (unused).
    """

    def __init__(self, orig_image_shape: tuple, crop_image_shape: tuple, calibration_parameter: dict = {},
                 solver_config: dict = {}, visualize_module=None):
        self.orig_image_shape = orig_image_shape
        self.crop_image_shape = crop_image_shape
        self.padding = solver_config["outer_padding"] if "outer_padding" in solver_config.keys() else 0
        self.pad_image_shape = (crop_image_shape[0] + self.padding, crop_image_shape[1] + self.padding)
        self.calib_param = calibration_parameter
        self.slv_config = solver_config
        self.visualizer = visualize_module
        self.setup_filter_preprocess()

        self._cuda_available = torch.cuda.is_available()
        self._device = "cuda" if self._cuda_available else "cpu"

        self.orig_imager = event_image_converter.EventImageConverter(self.orig_image_shape)
        self.crop_imager = event_image_converter.EventImageConverter(self.crop_image_shape, outer_padding=self.padding)
        self.normalize_t_in_batch = True  # displacement, not velocity (src/solver/base.py:98)
        self.orig_warper = warp.Warp(self.orig_image_shape, normalize_t=self.normalize_t_in_batch,
                                     calib_param=self.calib_param)
        self.crop_warper = warp.Warp(self.crop_image_shape, normalize_t=self.normalize_t_in_batch,
                                     calib_param=self.calib_param)
        self.previous_frame_best_estimation = None
        self.sequential_video_list: List[str] = list()
        self.evaluation_text_list: List[str] = list()
        logger.info(f"Configuration: \n    {self.slv_config}")

    def setup_filter_preprocess(self):
        """ROI crop set-up (src/solver/base.py:108-120).  Only the CROP filter -- the one the shipped
        config enables -- is applied; BAF/HOT filters are preprocessing outside this package."""
        if "filter" in self.slv_config and "xmin" in self.slv_config["filter"].get("parameters", {}):
            self.preproc_filter = True
            params = self.slv_config["filter"]["parameters"]
            self.crop_xmin, self.crop_xmax = params["xmin"], params["xmax"]
            self.crop_ymin, self.crop_ymax = params["ymin"], params["ymax"]
            extra = self.slv_config["filter"].get("filters") or []
            if extra:
                logger.warning(f"filters {extra} are not implemented in event_based_bos_b200; only CROP is applied")
        else:
            logger.info("No filtering process for events!")
            self.preproc_filter = False
            self.crop_xmin, self.crop_ymin = 0, 0
            self.crop_xmax, self.crop_ymax = self.orig_image_shape

    def preprocess(self, events: np.ndarray) -> Tuple[np.ndarray, float]:
        """Crop to the ROI; returns (events, time span of the un-filtered batch) (src/solver/base.py:123-139)."""
        num_orig = len(events)
        time_period = events[:, 2].max() - events[:, 2].min()
        if self.preproc_filter:
            events = utils.crop_event(events, self.crop_xmin, self.crop_xmax, self.crop_ymin, self.crop_ymax)
            logger.info(f"After preprocessng {len(events)} out of {num_orig}.")
        logger.info(f"Event stats: {len(events)} events, in {time_period} sec.")
        return events, time_period

    def estimate(self, events: np.ndarray, *args, **kwargs) -> np.ndarray:
        """Run the optimisation: [n,4] events -> best flow [2,H,W]."""
        raise NotImplementedError

    def calculate_flow_error(self, pred_disp: np.ndarray, gt_flow: np.ndarray, timescale: float = 1.0,
                             events: Optional[np.ndarray] = None, roi: Optional[dict] = None) -> dict:
        """EPE / AE / nPE of a [2,H,W] displacement against ground truth, restricted to pixels that
        saw events when `events` is given (src/solver/base.py:289-317)."""
        if events is not None:
            event_mask = self.orig_imager.create_eventmask(events)[:, roi["xmin"]:roi["xmax"], roi["ymin"]:roi["ymax"]]
        else:
            event_mask = None
        flow_error = utils.calculate_flow_error_numpy(gt_flow[None], pred_disp[None], event_mask=event_mask)
        logger.info(f"{flow_error = } for time period {timescale} sec.")
        return flow_error
