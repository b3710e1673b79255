"""Version of the pgsd_sph_b200 package; ``version`` mirrors the reference's pgsd.version.version
(/root/reference/pgsd/pgsd/version.py:12 -- the GSD 3.2.0 lineage whose file format v2 we write)."""
version = "3.2.0"
__version__ = version + "+b200.1"
__all__ = ['version']
