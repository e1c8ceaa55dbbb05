from torch_m3gnet_b200.data import MaterialGraphKey  # noqa: F401
