"""`extern/chamfer3D/dist_chamfer_3D.py` of the reference (chamfer_3DFunction / chamfer_3DDist, :50-120) over
`lnb_chamfer_forward/backward`: same call signature, same four return values (dist1, dist2, idx1, idx2)."""
import torch
from torch import nn
from torch.autograd import Function

from .._lib import lib, check, u32, vp


def _s():
    return vp(torch.cuda.current_stream().cuda_stream)


class chamfer_3DFunction(Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        if not (xyz1.is_cuda and xyz2.is_cuda):
            raise RuntimeError("chamfer_3DDist: GPU tensors only (as in the reference)")
        batchsize, n, dim = xyz1.size()
        assert dim == 3, "Wrong last dimension for the chamfer distance 's input! Check with .size()"
        _, m, dim = xyz2.size()
        assert dim == 3, "Wrong last dimension for the chamfer distance 's input! Check with .size()"
        xyz1, xyz2 = xyz1.float().contiguous(), xyz2.float().contiguous()
        dev = xyz1.device
        dist1 = torch.zeros(batchsize, n, device=dev)
        dist2 = torch.zeros(batchsize, m, device=dev)
        idx1 = torch.zeros(batchsize, n, dtype=torch.int32, device=dev)
        idx2 = torch.zeros(batchsize, m, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            check(lib.lnb_chamfer_forward(vp(xyz1.data_ptr()), vp(xyz2.data_ptr()), u32(batchsize), u32(n), u32(m),
                                          vp(dist1.data_ptr()), vp(dist2.data_ptr()), vp(idx1.data_ptr()),
                                          vp(idx2.data_ptr()), _s()), "chamfer_forward")
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        ctx.mark_non_differentiable(idx1, idx2)
        return dist1, dist2, idx1, idx2

    @staticmethod
    def backward(ctx, graddist1, graddist2, gradidx1, gradidx2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        graddist1, graddist2 = graddist1.contiguous(), graddist2.contiguous()
        gradxyz1, gradxyz2 = torch.zeros_like(xyz1), torch.zeros_like(xyz2)
        b, n, _ = xyz1.shape
        m = xyz2.shape[1]
        with torch.cuda.device(xyz1.device):
            check(lib.lnb_chamfer_backward(vp(xyz1.data_ptr()), vp(xyz2.data_ptr()), vp(gradxyz1.data_ptr()),
                                           vp(gradxyz2.data_ptr()), vp(graddist1.data_ptr()), vp(graddist2.data_ptr()),
                                           vp(idx1.data_ptr()), vp(idx2.data_ptr()), u32(b), u32(n), u32(m), _s()),
                  "chamfer_backward")
        return gradxyz1, gradxyz2


class chamfer_3DDist(nn.Module):
    def forward(self, input1, input2):
        return chamfer_3DFunction.apply(input1.contiguous(), input2.contiguous())
