"""Drop-in import path of the reference's model/point_sdf_net.py: the B200-native classes."""
from shapegan_b200.nn.point_sdf_net import PointNet, SDFGenerator, scatter_max  # noqa: F401
