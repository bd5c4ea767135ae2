"""Unit handling of the reference (gala UnitSystem + astropy G, main.py:18-35) without those packages.

The reference fixes kpc / Myr / Msun / rad everywhere (`usys`, main.py:18) and takes
`G = astropy.constants.G.decompose(usys)`; the value is pinned to 15 digits by the reference's own
notebook (examples/custom_potential.ipynb cell 6).
"""
G_KPC_MYR_MSUN = 4.498502151469553e-12


class UnitSystem:
    def __init__(self, *names, G=None):
        self.names = tuple(str(n) for n in names)
        self.G = G

    def __repr__(self):
        return "UnitSystem(" + ", ".join(self.names) + ")"


usys = UnitSystem("kpc", "Myr", "Msun", "rad", G=G_KPC_MYR_MSUN)
dimensionless = UnitSystem(G=1.0)


def resolve_G(units):
    """`Potential.__init__` (main.py:22-30): units None -> dimensionless (G = 1), else G in those units."""
    if units is None or units is dimensionless:
        return 1.0
    if isinstance(units, UnitSystem):
        if units.G is None:
            raise ValueError("UnitSystem without G")
        return units.G
    # a real gala UnitSystem (if the user has gala/astropy installed)
    try:
        from astropy.constants import G as _G
        return float(_G.decompose(units).value)
    except Exception as exc:  # pragma: no cover - astropy is not in the build image
        raise TypeError(f"cannot derive G from units={units!r}; pass streamsculptor_b200.usys") from exc
