"""Stand-in for timm: deit_models.py:10-11 imports it; only load_pretrained_weights would call it."""
