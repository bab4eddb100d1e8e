"""lightly_train_b200: B200-native (sm_100a) DINOv2 training step behind lightly-train's method API."""
__version__ = "0.1.0"
