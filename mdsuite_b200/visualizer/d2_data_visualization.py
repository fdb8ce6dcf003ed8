"""
2-D line plots written to ``<experiment>/figures/<title>.html``
(reference mdsuite/visualizer/d2_data_visualization.py:39-115 uses bokeh; bokeh is used here too
when it is importable, otherwise a self-contained HTML/SVG page with the same file name is written).
"""
from __future__ import annotations

import html
import math
import os
from typing import List

from mdsuite_b200.utils.config import config


class DataVisualizer2D:
    def __init__(self, title: str, path: str = "."):
        self.title = title
        self.path = str(path)
        os.makedirs(self.path, exist_ok=True)
        self.output_file = os.path.join(self.path, f"{title}.html")
        try:
            import bokeh  # noqa: F401
            self._bokeh = True
        except ImportError:
            self._bokeh = False

    def construct_plot(self, x_data, y_data, x_label: str, y_label: str, title: str,
                       layouts: list = None):
        if self._bokeh:
            from bokeh.plotting import figure
            fig = figure(x_axis_label=x_label, y_axis_label=y_label, title=title)
            fig.line(x_data, y_data)
            return fig
        return {"x": list(x_data), "y": list(y_data), "x_label": x_label, "y_label": y_label,
                "title": title}

    def grid_show(self, figures: List):
        if self._bokeh:
            from bokeh.io import output_file, save, show
            from bokeh.layouts import gridplot
            output_file(self.output_file, title=self.title)
            grid = gridplot(figures, ncols=3, sizing_mode=config.bokeh_sizing_mode)
            (show if config.show_plots else save)(grid)
            return self.output_file
        with open(self.output_file, "w") as fh:
            fh.write(self._render_html(figures))
        return self.output_file

    # ---- dependency-free fallback renderer --------------------------------------------------
    @staticmethod
    def _svg(fig, w=420, h=300, pad=45) -> str:
        pts = [(x, y) for x, y in zip(fig["x"], fig["y"])
               if isinstance(y, (int, float)) and math.isfinite(y) and math.isfinite(x)]
        if not pts:
            return f"<svg width='{w}' height='{h}'></svg>"
        xs, ys = [p[0] for p in pts], [p[1] for p in pts]
        x0, x1, y0, y1 = min(xs), max(xs), min(0.0, min(ys)), max(ys)
        sx = (w - 2 * pad) / ((x1 - x0) or 1.0)
        sy = (h - 2 * pad) / ((y1 - y0) or 1.0)
        path = " ".join(f"{'M' if i == 0 else 'L'}{pad + (x - x0) * sx:.1f},{h - pad - (y - y0) * sy:.1f}"
                        for i, (x, y) in enumerate(pts))
        return (f"<svg width='{w}' height='{h}' xmlns='http://www.w3.org/2000/svg'>"
                f"<rect x='{pad}' y='{pad}' width='{w - 2 * pad}' height='{h - 2 * pad}' fill='none' stroke='#888'/>"
                f"<path d='{path}' fill='none' stroke='#1f77b4' stroke-width='1.5'/>"
                f"<text x='{w / 2}' y='20' text-anchor='middle' font-size='14'>{html.escape(fig['title'])}</text>"
                f"<text x='{w / 2}' y='{h - 8}' text-anchor='middle' font-size='11'>{html.escape(fig['x_label'])}"
                f"  [{x0:.3g}, {x1:.3g}]</text>"
                f"<text x='12' y='{h / 2}' font-size='11' transform='rotate(-90 12 {h / 2})' "
                f"text-anchor='middle'>{html.escape(fig['y_label'])}  [{y0:.3g}, {y1:.3g}]</text></svg>")

    def _render_html(self, figures) -> str:
        cells = "".join(f"<div style='display:inline-block;margin:6px'>{self._svg(f)}</div>"
                        for f in figures)
        return (f"<!DOCTYPE html><html><head><meta charset='utf-8'><title>{html.escape(self.title)}"
                f"</title></head><body><h3>{html.escape(self.title)}</h3>{cells}</body></html>")
