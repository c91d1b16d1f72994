"""B200-native batched caption decoder (beam / greedy / sampling) behind the
zyj0021200/simpleImageCaptionZoo Engine API.  See DESIGN.md."""
__version__ = "0.1.0"
