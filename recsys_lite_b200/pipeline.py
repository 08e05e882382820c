"""``RecsysPipeline`` of the reference (tools.py:222-266) with ``recommend`` routed to the fused device path.

fit / predict chain the steps exactly as the reference does (same error wrapping, same shape checks).  ``recommend``:
when the LAST step is a fitted ``RecommenderSystem`` of this package and no earlier step transforms the data, the call
goes to its fused ``recommend`` (scores never materialised); otherwise it is the reference's ``top_n(predict(X), X)``
with both stages on the device.
"""
from __future__ import annotations

from typing import Any, List

import numpy as np

from .api import RecommenderSystem, top_n


class RecsysPipeline:
    def __init__(self, steps: List[tuple]):
        if not steps:
            raise ValueError("Pipeline must have at least one step")
        self.steps = steps

    def fit(self, X, **fit_params) -> "RecsysPipeline":
        data = X
        for name, step in self.steps:
            try:
                if hasattr(step, "fit"):
                    step.fit(data, **fit_params)
                if hasattr(step, "transform"):
                    new_data = step.transform(data)
                    if new_data.shape != data.shape:
                        raise ValueError(f"Shape mismatch after step '{name}': {new_data.shape} vs {data.shape}")
                    data = new_data
            except Exception as e:
                raise RuntimeError(f"Error in pipeline step '{name}': {str(e)}") from e
        return self

    def predict(self, X):
        data = X
        for name, step in self.steps:
            try:
                if hasattr(step, "predict"):
                    new_data = step.predict(data)
                elif hasattr(step, "transform"):
                    new_data = step.transform(data)
                else:
                    raise AttributeError(f"Step '{name}' lacks predict or transform")
                if new_data.shape != data.shape:
                    raise ValueError(f"Shape mismatch after step '{name}': {new_data.shape} vs {data.shape}")
                data = new_data
            except Exception as e:
                raise RuntimeError(f"Error in pipeline step '{name}': {str(e)}") from e
        return data

    def recommend(self, X, n: int = 10) -> list:
        last: Any = self.steps[-1][1]
        passthrough = all(not hasattr(step, "transform") and not hasattr(step, "predict") for _, step in self.steps[:-1])
        if isinstance(last, RecommenderSystem) and last.is_fitted() and last.algorithm in ("als", "svd") and passthrough:
            return np.asarray(last.recommend(X, n=n)).tolist()
        preds = self.predict(X)
        return top_n(preds, X, n=n).tolist()
