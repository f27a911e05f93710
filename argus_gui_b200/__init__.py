"""argus_gui_b200 - B200-native backend for Argus' calibration / reconstruction hot path.

Sub-modules mirror the reference's module names for the path (argus_gui/sbaDriver.py, triangulate.py, tools.py) plus the
``sba`` module python-sba provides to Argus.  Nothing here imports a GUI toolkit."""
__version__ = "0.1.0"
