"""Mirror of the names/values the hot path reads from /root/reference/src/config/gan_config.py:1-21.
Module-level constants, mutated in place by callers and read at CALL time, like the reference."""
epochs: int = 500
batch_size: int = 128      # GLOBAL batch; a data-parallel rank processes batch_size // world_size rows

z_size: int = 128

# NOT in the reference (its layer widths are hard-coded formulas, cvae_gan_models.py:16-18,85-87,173-175,257-259): None keeps
# them; (h1, h2, h3) builds the widened model of BASELINE.json configs[4] - e.g. (1024, 512, 256) - with these three hidden
# widths in all four networks (multiples of 64, <= 1024, h2 <= 512)
hidden = None

g_lr: float = 2e-4
g_loop_num: int = 3

d_lr: float = 2e-4
d_loop_num: int = 5

c_lr: float = 1e-4
c_loop_num: int = 5

cvae_gan_config = {
    'lambda_recon': 1.0,
    'lambda_kl': 0.1,
    'lambda_adv': 1.0,
    'lambda_class': 0.5,
    'confidence_threshold': 0.5,
}

# sibling trainer CGAN (/root/reference/src/config/gan_config.py:40-44, read by src/cgan.py:41-42,165-171,262)
cgan_config = {
    'lambda_adv': 1.0,
    'lambda_class': 0.5,
    'confidence_threshold': 0.5,
}

# sibling trainer CVAE (/root/reference/src/config/gan_config.py:51-56, read by src/cvae.py:40-42,145-151,274)
cvae_config = {
    'lambda_recon': 1.0,
    'lambda_kl': 0.01,
    'lambda_class': 0.1,
    'confidence_threshold': 0.5,
}

# sibling trainer VAE-GAN (/root/reference/src/config/gan_config.py:33-38, read by src/vae_gan.py:29-31,131-135)
vae_gan_config = {
    'lambda_recon': 1.0,
    'lambda_kl': 0.01,
    'lambda_adv': 0.1,
    'confidence_threshold': 0.5,
}
