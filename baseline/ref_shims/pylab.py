"""Empty stand-in for pylab (base_model.py:3 imports it, nothing uses it)."""
