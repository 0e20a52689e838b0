"""b200-smplx-fit: B200-native SMPL-X fitting engine (per-frame SMPLify-X inner loop).

Host-side mirror of the reference's ``smplifyx`` Python surface over a C-ABI CUDA library
(``include/sfx.h``).  Import as ``smplifyx_b200``.
"""
__version__ = '0.1.0'
