import numpy as np
import torch


class ThreeDEvaluator:
    r"""Mean-absolute-error evaluator with the interface of reference
    dig/threedgraph/evaluation/eval.py:4-33 (`eval({'y_true','y_pred'}) -> {'mae': float}`)."""

    def eval(self, input_dict):
        if 'y_pred' not in input_dict or 'y_true' not in input_dict:
            raise AssertionError("input_dict needs 'y_true' and 'y_pred'")
        y_pred, y_true = input_dict['y_pred'], input_dict['y_true']
        both_np = isinstance(y_true, np.ndarray) and isinstance(y_pred, np.ndarray)
        both_t = isinstance(y_true, torch.Tensor) and isinstance(y_pred, torch.Tensor)
        assert both_np or both_t, "y_true and y_pred must both be numpy arrays or both torch tensors"
        assert y_true.shape == y_pred.shape
        if both_t:
            return {'mae': (y_pred - y_true).abs().mean().cpu().item()}
        return {'mae': float(np.abs(y_pred - y_true).mean())}
