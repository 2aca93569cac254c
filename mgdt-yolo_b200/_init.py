"""mgdt-yolo_b200: B200-native (sm_100a) detection forward path of MGDT-YOLO.

Only cheap, CPU-safe names are imported here; the CUDA library is loaded lazily
by ``_lib.lib()`` the first time a kernel is needed and fails loudly if absent.
"""
__version__ = "0.1.0"
