"""
Physical constants and unit conversions used by the prism forward model.

Same names and values as the reference's fatiando/constants.py:19-31 (only the
ones the prism path uses, prism.py:91).  The unit products (``G*SI2MGAL`` ...)
are formed in Python exactly where the reference forms them, so they carry the
same rounding.
"""

#: Conversion factor from SI units to Eotvos: 1/s**2 = 10**9 Eotvos
SI2EOTVOS = 1000000000.0

#: Conversion factor from SI units to mGal: 1 m/s**2 = 10**5 mGal
SI2MGAL = 100000.0

#: The gravitational constant in m**3 kg**-1 s**-1
G = 0.00000000006673

#: Proportionality constant used in the magnetic method in henry/m (SI)
CM = 10. ** (-7)

#: Conversion factor from tesla to nanotesla
T2NT = 10. ** (9)
