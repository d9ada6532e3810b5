"""The slice of the reference's data classes the hot path touches (src/NanoRepeat/repeat_region.py:32-55,
:116-151).  Attribute names are the reference's, so the reference's own Read / RepeatRegion objects can be
passed to nanorepeat_b200.round1_and_round2_estimation / round3_estimation unchanged (duck typing)."""


class Read:
    def __init__(self, read_name=None, dist_between_anchors=None):
        self.read_name = read_name
        self.strand = None
        self.dist_between_anchors = dist_between_anchors
        self.round1_repeat_size = None
        self.round2_repeat_size = None
        self.round3_repeat_size = None
        self.round3_paf_text = ""     # kept for attribute parity; the binary path never fills it


class RepeatRegion:
    def __init__(self):
        self.left_anchor_seq = None
        self.right_anchor_seq = None
        self.left_anchor_len = None
        self.right_anchor_len = None
        self.repeat_unit_seq = None
        self.read_dict = dict()            # read_name -> Read   (insertion order = processing order)
        self.read_core_seq_dict = dict()   # read_name -> core sequence
        self.temp_out_dir = None
        self.temp_file_list = []

    @classmethod
    def from_synth(cls, reg):
        """Build from nanorepeat_b200.synth.SynthRegion."""
        rr = cls()
        rr.left_anchor_seq = reg.left_anchor_seq
        rr.right_anchor_seq = reg.right_anchor_seq
        rr.left_anchor_len = len(reg.left_anchor_seq)
        rr.right_anchor_len = len(reg.right_anchor_seq)
        rr.repeat_unit_seq = reg.repeat_unit_seq
        for name, core, dist in zip(reg.read_names, reg.core_seqs, reg.dist_between_anchors):
            rr.read_dict[name] = Read(name, dist)
            rr.read_core_seq_dict[name] = core
        return rr
