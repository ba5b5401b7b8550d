"""Mirror of /root/reference/src/config/classifier_config.py:1-3."""
epochs: int = 100
lr: float = 1e-3
batch_size: int = 64
