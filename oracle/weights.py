"""TEST INFRASTRUCTURE ONLY — re-export of the seeded synthetic-weight generator.

The generator itself lives in deepv_b200/synthetic.py (it contains no reference arithmetic: it
only enumerates parameter names/shapes and draws seeded normals), so that bench.py's GPU arm
does not have to import anything from oracle/.
"""
from deepv_b200.synthetic import (MMDIT_DEFAULT, VAE_DEFAULT, make_weights, mmdit_shapes,  # noqa: F401
                                  mmdit_weights, vae_decoder_shapes, vae_encoder_shapes, vae_weights)
