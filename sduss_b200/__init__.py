"""sduss_b200: B200-native (sm_100a) drop-in for the mixed-resolution denoising step of
MiRaCLeXeoN/sduss (Mixfusion). See DESIGN.md for scope and INTEGRATION.md for the binding."""
__version__ = "0.1.0"
