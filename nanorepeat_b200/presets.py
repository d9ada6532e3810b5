"""Per-data-type scoring presets (reference src/NanoRepeat/tk.py:502-517).

At the surveyed reference version every data type maps to minimap2 `-x map-ont`, so the five rows carry the
same scoring; the table is kept so the CLI's -d option keeps its meaning.  Values come from the C ABI
(nr_get_preset) so Python and C agree by construction.
"""
import sys

DATA_TYPES = ("ont", "ont_sup", "ont_q20", "clr", "hifi")


def get_preset_for_minimap2(data_type):
    """Same contract as tk.get_preset_for_minimap2: the minimap2 preset string, exit(1) on unknown type."""
    if data_type in DATA_TYPES:
        return " -x map-ont "
    sys.stderr.write(f"ERROR: Unknown data type: {data_type}\n\n")
    sys.exit(1)


def get_scoring(data_type):
    """nr_scoring_t for a data type (raises ValueError on unknown type)."""
    from . import engine
    return engine.get_preset(data_type)
