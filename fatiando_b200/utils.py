"""
Direction-cosine helpers of the prism path (reference: fatiando/utils.py).
"""
import numpy


def dircos(inc, dec):
    """
    Unit vector of a direction given by inclination and declination (degrees).

    x -> North, y -> East, z -> Down; inclination positive down, declination
    measured from x.  Same expression order as fatiando/utils.py:322-348, so
    the three cosines are bit-identical to the reference's.
    """
    d2r = numpy.pi / 180.
    ci = numpy.cos(d2r * inc)
    return [ci * numpy.cos(d2r * dec), ci * numpy.sin(d2r * dec),
            numpy.sin(d2r * inc)]


def ang2vec(intensity, inc, dec):
    """
    Vector from intensity, inclination and declination (degrees).

    ``intensity`` may be an array; then one row per value is returned
    (fatiando/utils.py:285-319).
    """
    return numpy.transpose([intensity * c for c in dircos(inc, dec)])


def vec2ang(vector):
    """
    Intensity, inclination and declination (degrees) of a 3-vector
    (fatiando/utils.py:252-282).
    """
    x, y, z = vector
    intensity = numpy.linalg.norm(vector)
    r2d = 180. / numpy.pi
    return [intensity, r2d * numpy.arcsin(z / intensity), r2d * numpy.arctan2(y, x)]


def gaussian2d(x, y, sigma_x, sigma_y, x0=0, y0=0, angle=0.0):
    """
    Non-normalised 2-D Gaussian bump (fatiando/utils.py:530-567); used to build
    the synthetic topography of the relief benchmark.
    """
    theta = -1 * angle * numpy.pi / 180.
    tmpx = 1. / sigma_x ** 2
    tmpy = 1. / sigma_y ** 2
    sintheta = numpy.sin(theta)
    costheta = numpy.cos(theta)
    a = tmpx * costheta + tmpy * sintheta ** 2
    b = (tmpy - tmpx) * costheta * sintheta
    c = tmpx * sintheta ** 2 + tmpy * costheta ** 2
    xhat = x - x0
    yhat = y - y0
    return numpy.exp(-(a * xhat ** 2 + 2. * b * xhat * yhat + c * yhat ** 2))
