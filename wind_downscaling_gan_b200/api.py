"""Placeholder filled in below."""
from .gan.models import make_generator, make_discriminator  # noqa: F401
