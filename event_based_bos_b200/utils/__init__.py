"""Utilities on the contrast-maximisation path: synthetic inputs (src/utils/event_utils.py:18-47,
src/utils/flow_utils.py:20-45), event cropping (src/utils/event_utils.py:109-129) and the flow-error
metrics (src/utils/flow_utils.py:769-821)."""
from .event_utils import crop_event, generate_events, rebase_time, synthetic_bos_events, synthetic_events
from .flow_utils import (calculate_flow_error_numpy, calculate_flow_error_tensor, generate_dense_optical_flow, generate_uniform_optical_flow,
                         smooth_flow, synthetic_flow)
