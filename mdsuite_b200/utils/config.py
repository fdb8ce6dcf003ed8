"""Global knobs (reference mdsuite/utils/config.py:31-59)."""
from dataclasses import dataclass


@dataclass
class Config:
    jupyter: bool = False
    bokeh_sizing_mode: str = "stretch_both"
    memory_fraction: float = 0.5  # of free HOST memory for pinned staging; of free HBM for frames
    show_plots: bool = False  # the reference opens a browser tab; off by default on a GPU node


config = Config()
