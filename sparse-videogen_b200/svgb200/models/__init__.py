"""Per-model host logic mirroring svg/models/<m>/{attention,utils,inference}.py (the attention core only)."""
