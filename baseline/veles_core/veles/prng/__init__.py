"""``veles.prng``: seedable host generators addressed by key (``prng.get(2)``)."""
from veles.prng.random_generator import RandomGenerator, get  # noqa: F401
from veles.prng import uniform  # noqa: F401
