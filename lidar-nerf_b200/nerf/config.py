"""Configuration of the fused LiDAR-field step (no torch, no ctypes: importable by the CPU reference arm of bench.py
and by the oracle without loading liblnb200.so).  Defaults = BASELINE.json configs[1]
(main_lidarnerf.py defaults + configs/kitti360_1908.txt)."""
import math
from dataclasses import dataclass


@dataclass
class FieldConfig:
    # scene / march (main_lidarnerf.py defaults; configs/kitti360_1908.txt)
    bound: float = 1.0
    grid_size: int = 128                 # renderer.py:75
    min_near_lidar: float = 0.010784853507573345   # = opt.scale (main_lidarnerf.py:286-287)
    far_factor: float = 81.0             # renderer.py:134-138
    dt_gamma: float = 0.0
    max_steps: int = 1024
    T_thresh: float = 1e-4
    density_scale: float = 1.0
    density_thresh: float = 10.0         # main_lidarnerf.py:210-215
    # hash grid (configs/kitti360_1908.txt:7, main_lidarnerf.py:68-69)
    num_levels: int = 16
    level_dim: int = 2
    base_resolution: int = 16
    desired_resolution: int = 32768
    log2_hashmap_size: int = 19
    # MLPs (ffmlp 64x2 each; network.py:45-99 shapes)
    mlp_dtype: str = "fp16"              # "fp16" | "bf16": element type of the MLP weights / activations / activation
                                         # gradients on the tensor cores (fp32 accumulation either way; the hash table and
                                         # its gradient path stay fp16 / fp32).  bf16 = BASELINE config 5; needs the
                                         # persistent forward kernel (fused_gather)
    hidden_dim: int = 64
    sigma_layers: int = 2                # FFMLP num_layers
    head_layers: int = 2
    freq_degree: int = 12                # network.py:83
    dir_encoding: str = "frequency"      # "frequency" (LiDAR head, network.py:83) | "sh" (spherical harmonics of degree
    sh_degree: int = 4                   #   sh_degree, the direction encoder of network.py:64 / network_tcnn.py:74-80)
    geo_feat_dim: int = 15
    # loss (configs/kitti360_1908.txt:2-4)
    alpha_d: float = 1e3
    alpha_r: float = 1.0
    alpha_i: float = 10.0
    # patch depth-gradient loss (configs/kitti360_1908.txt:5-6,8: alpha_grad = 100, grad_loss = True, patches of 2 x 8
    # pixels every other epoch; nerf/utils.py:748-876).  patch = (1, 1) or alpha_grad = 0: per-ray loss only.
    patch_size: tuple = (1, 1)
    alpha_grad: float = 0.0
    grad_clip: float = 0.01              # nerf/utils.py:846-849
    # optimiser (main_lidarnerf.py:389-391, lr default 1e-2)
    lr: float = 1e-2
    beta1: float = 0.9
    beta2: float = 0.99
    eps: float = 1e-15
    loss_scale: float = 128.0            # static loss scale for the fp16 gradient chain (GradScaler's role)
    grid_update_interval: int = 16
    perturb: bool = True                 # jitter the march start (Trainer.train_step passes perturb=True)
    fused_field: bool = True             # density MLP + LiDAR head as the fused field kernels (csrc/field.cu)
    fused_gather: bool = True            # hash-grid gather + density MLP + LiDAR head forward as ONE persistent kernel
                                         # (csrc/field_fused.cu); needs fused_field
    l2_persist_table: bool = True        # persistent forward kernel: access-policy window keeping the fp16 table in L2
    fused_composite: bool = True         # composite fwd + LiDAR loss + composite bwd as one kernel (csrc/raymarching.cu)
    compact_backward: bool = True        # backward kernels walk only the samples up to each ray's early stop
    late_grad_zero: bool = True          # zero the gradient table right before the scatter (L2-resident) instead of in Adam
    fused_exchange: bool = True          # data parallel: one peer-memory kernel (reduce-scatter + Adam + all-gather) over
                                         # NVLink via torch symmetric memory; falls back to NCCL when it cannot be set up
    multicast_exchange: bool = False     # ... through NVSwitch multicast / in-switch reduction (multimem.ld_reduce / .st, NVLS)
                                         # instead of peer loads/stores.  Correct (2-GPU test) but measured SLOWER on B200:
                                         # 0.741 vs 0.651 ms/step at 2 GPUs, 0.849 vs 0.812 at 8 (profiles/r02_summary.md)
    overlap_exchange: bool = True        # data parallel, graph mode: the exchange overlaps the next step's march
    pipeline_adam: bool = False          # graph mode, one rank: Adam of step i runs next to the march of step i+1
                                         # (measured: -6 us/step device time, +CPU launch work; off by default)
    seed: int = 0

    @property
    def cascade(self):
        return 1 + math.ceil(math.log2(self.bound))   # renderer.py:74

    @property
    def patch_loss(self):
        return self.alpha_grad > 0 and self.patch_size[0] * self.patch_size[1] > 1

    @property
    def dir_dim(self):
        """Columns of the head input taken by the direction encoding."""
        return self.sh_degree ** 2 if self.dir_encoding == "sh" else 3 + 6 * self.freq_degree

    @property
    def dir_code(self):
        """The `degree` argument of the lnb_field_* entry points (LNB_DIR_SH(deg) = 0x100 | deg for SH)."""
        return (0x100 | self.sh_degree) if self.dir_encoding == "sh" else self.freq_degree

    @property
    def head_in_dim(self):
        raw = self.dir_dim + self.geo_feat_dim                     # 75 + 15 = 90 (frequency 12) / 16 + 15 = 31 (SH 4)
        return (raw + 15) // 16 * 16                               # padded to 96 / 32 for the tensor cores
