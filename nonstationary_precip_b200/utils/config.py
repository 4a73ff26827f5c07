"""Constants of the reference's utils/config.py:10-20 (kept for import compatibility)."""
from pathlib import Path

import torch

TORCH_VERSION = torch.__version__
AVAILABLE_GPU = torch.cuda.device_count()
GPU_ACTIVE = bool(AVAILABLE_GPU)
EPSILON = 1e-5
BASE_SEED = 173
BASE_PATH = Path(__file__).parent.parent.parent
RESULTS_DIR = BASE_PATH / "results"
DATASET_DIR = BASE_PATH / "data"
