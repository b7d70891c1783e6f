"""`lidarnerf.nerf.network_tcnn.NeRFNetwork` without tiny-cuda-nn.

`main_lidarnerf.py:289-308` imports this module path when `--tcnn` / `-L` is given and passes the keyword set of the
reference's tcnn network (network_tcnn.py:10-28: `encoding="HashGrid"`, `n_features_per_level`, ...).  The reference
module needs `tinycudann`, an un-vendored, unpinned dependency that cannot be installed offline, so
this class offers the same constructor / `density` / `color` / `get_params` contract on the in-tree building blocks of
this library (GridEncoder + FFMLP + SH / frequency encoders).  What it keeps from the tcnn variant: the resolution is
scaled by `bound` (network_tcnn.py:40-42), `n_features_per_level` is honoured, the MLP depths map
`n_hidden_layers = num_layers - 1`.  What it cannot keep: tiny-cuda-nn's own table sizing / hashing and its Frequency
encoding convention - PARITY AT THE TCNN BOUNDARY IS UNPINNED; checkpoints of the two are not
interchangeable, exactly as the reference's two variants are not interchangeable with each other.

`compat.install()` registers this module as `lidarnerf.nerf.network_tcnn`, so the unmodified entry script resolves
the import without tinycudann.
"""
from .network import NeRFNetwork as _InTreeNetwork


class NeRFNetwork(_InTreeNetwork):
    def __init__(self, encoding="HashGrid", desired_resolution=2048, log2_hashmap_size=19,
                 encoding_dir="SphericalHarmonics", n_features_per_level=2, num_layers=2, hidden_dim=64, geo_feat_dim=15,
                 num_layers_color=3, hidden_dim_color=64, out_color_dim=3, out_lidar_color_dim=2, bound=1, **kwargs):
        names = {"HashGrid": "hashgrid", "hashgrid": "hashgrid", "TiledGrid": "tiledgrid", "tiledgrid": "tiledgrid",
                 "Frequency": "frequency", "frequency": "frequency"}
        if encoding not in names:
            raise NotImplementedError(f"encoding {encoding!r}: choose from {sorted(names)}")
        self.n_features_per_level = n_features_per_level
        self.desired_resolution = desired_resolution
        self.log2_hashmap_size = log2_hashmap_size
        super().__init__(encoding=names[encoding], encoding_dir="sphere_harmonics",
                         desired_resolution=int(desired_resolution * bound),          # network_tcnn.py:40-42
                         log2_hashmap_size=log2_hashmap_size, num_layers=num_layers, hidden_dim=hidden_dim,
                         geo_feat_dim=geo_feat_dim, num_layers_color=num_layers_color,
                         hidden_dim_color=hidden_dim_color, out_color_dim=out_color_dim,
                         out_lidar_color_dim=out_lidar_color_dim, bound=bound, use_ffmlp=True,
                         level_dim=n_features_per_level, **kwargs)
