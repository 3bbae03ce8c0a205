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


class SparseList(object):
    """
    A fixed-length list that stores only the entries that were set; everything
    else reads as 0.0 (the container of the harvester's estimates,
    utils.py:351-430 of the reference).

    >>> l = SparseList(5)
    >>> l[3] = 42.0
    >>> len(l), l[1], l[3]
    (5, 0.0, 42.0)
    >>> l[1] += 3.0
    >>> list(l)
    [0.0, 3.0, 0.0, 42.0, 0.0]
    """

    def __init__(self, size, elements=None):
        self.size = size
        self.elements = {} if elements is None else elements

    def __str__(self):
        return str(self.elements)

    def __len__(self):
        return self.size

    def __iter__(self):
        return (self[i] for i in range(self.size))

    def _position(self, index):
        if index < 0:
            index += self.size
        if index < 0 or index >= self.size:
            raise IndexError('index out of range')
        return index

    def __getitem__(self, index):
        return self.elements.get(self._position(index), 0.0)

    def __setitem__(self, index, value):
        if index >= self.size:
            raise IndexError('index out of range')
        self.elements[index] = value
