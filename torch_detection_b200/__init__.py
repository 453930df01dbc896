"""B200-native ResNet + FPN feature extraction, drop-in behind Torch_Detection's build API."""
