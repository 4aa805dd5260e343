"""dss_b200 -- B200-native (sm_100a) differentiable surface splatting rasterizer: the hot path of
yifita/DSS (DSS/csrc + external/prefix_sum + FRNN radius binning + the glue of DSS/core/{rasterizer,
renderer}.py) behind the reference's own operator API.  See DESIGN.md and INTEGRATION.md."""
__version__ = "0.1.0"
