"""FFT feed of the reference's plot consumers (src/plots/*), on the device.  The Qt widgets are not
part of this build; ``feeds`` computes what their ``update`` methods draw."""
from .feeds import SpectrumFeed, WaterfallFeed, powerSpectrum, stftDb  # noqa: F401
