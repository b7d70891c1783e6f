"""`extern/fscore.py` of the reference: F-score of two point clouds from their (squared) Chamfer distances."""
import torch


def fscore(dist1, dist2, threshold=0.001):
    precision_1 = torch.mean((dist1 < threshold).float(), dim=1)
    precision_2 = torch.mean((dist2 < threshold).float(), dim=1)
    f = 2 * precision_1 * precision_2 / (precision_1 + precision_2)
    f[torch.isnan(f)] = 0
    return f, precision_1, precision_2
