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


def contaminate(data, stddev, percent=False, return_stddev=False, seed=None):
    """
    Add pseudorandom, zero-mean gaussian noise to ``data`` (an array, or a list of arrays
    with a list of ``stddev``), like the reference's ``utils.contaminate``
    (utils.py:420-505): the noise vector has its mean removed; ``percent=True`` takes
    ``stddev`` as a fraction of ``max(abs(data))``; the same ``seed`` gives the same noise as
    the reference (numpy's legacy generator).

    >>> import numpy as np
    >>> noisy = contaminate(np.ones(5), 0.1, seed=0)
    >>> print(np.array2string(noisy, precision=8))
    [1.03137726 0.89498775 0.95284582 1.07906135 1.04172782]
    >>> noisy, std = contaminate(np.ones(5), 0.05, seed=0, percent=True, return_stddev=True)
    >>> print(std)
    0.05
    """
    rng = numpy.random.RandomState(seed)
    single = not isinstance(stddev, list)
    stddevs = [stddev] if single else list(stddev)
    arrays = [data] if single else list(data)
    contam = []
    for i, arr in enumerate(arrays):
        if stddevs[i] == 0.:
            contam.append(arr)
            continue
        if percent:
            stddevs[i] = stddevs[i] * max(abs(arr))
        noise = rng.normal(scale=stddevs[i], size=len(arr))
        noise -= noise.mean()
        contam.append(numpy.array(arr) + noise)
    if single:
        contam, stddevs = contam[0], stddevs[0]
    return [contam, stddevs] if return_stddev else contam


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
