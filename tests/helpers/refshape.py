"""A module shaped like the slice of the reference's NanoRepeat.nanoRepeat_bam that the drop-in patches (written for
these tests, not copied: /root/reference does not exist on the GPU box).  What matters is the SHAPE the reference has:

  * module-level functions round1_and_round2_estimation(data_type, repeat_region, num_cpu) and
    round3_estimation(data_type, fast_mode, repeat_region, num_cpu), looked up through the module's globals by the caller
    (nanoRepeat_bam.py:675-679), so that install(module) takes effect;
  * plain data classes Read / RepeatRegion with the reference's attribute names (repeat_region.py:32-55, :116-151),
    including the ones the hot path never touches;
  * a worker that runs regions i = pid, pid + P, ... and returns them through a multiprocessing queue, i.e. PICKLED
    (nanoRepeat_bam.py:602-612), after being forked from a parent that has already imported everything (:719-724).
"""


class Read:
    def __init__(self):
        self.read_name = None
        self.full_read_len = None
        self.core_seq = None
        self.init_repeat_size = None
        self.left_anchor_is_good = False
        self.right_anchor_is_good = False
        self.both_anchors_are_good = False
        self.core_seq_start_pos = None
        self.core_seq_end_pos = None
        self.mid_seq_start_pos = None
        self.mid_seq_end_pos = None
        self.dist_between_anchors = None
        self.seq_between_anchors = None
        self.left_buffer_len = None
        self.right_buffer_len = None
        self.strand = None
        self.round1_repeat_size = None
        self.round2_repeat_size = None
        self.round3_repeat_size = None
        self.round3_paf_text = ""


class RepeatRegion:
    def __init__(self):
        self.left_anchor_seq = None
        self.right_anchor_seq = None
        self.left_anchor_len = None
        self.right_anchor_len = None
        self.repeat_unit_seq = None
        self.chrom = "chr4"
        self.start_pos = 0
        self.end_pos = 0
        self.region_fq_file = None
        self.core_seq_fq_file = None
        self.temp_out_dir = None
        self.temp_file_list = []
        self.read_dict = dict()
        self.read_core_seq_dict = dict()
        self.results = None

    def to_unique_id(self):
        return f"{self.chrom}-{self.start_pos}-{self.end_pos}-{self.repeat_unit_seq}"


def round1_and_round2_estimation(data_type, repeat_region, num_cpu):
    raise RuntimeError("the reference's own round1_and_round2_estimation would shell out to pyminimap2 here")


def round3_estimation(data_type, fast_mode, repeat_region, num_cpu):
    raise RuntimeError("the reference's own round3_estimation would shell out to pyminimap2 here")


def quantify1repeat(process_name, num_threads_per_region, fast_mode, data_type, repeat_region):
    # the two calls of quantify1repeat_from_bam (nanoRepeat_bam.py:675-679), through the module's globals
    round1_and_round2_estimation(data_type, repeat_region, num_threads_per_region)
    round3_estimation(data_type, fast_mode, repeat_region, num_threads_per_region)
    return repeat_region


def worker(process_id, num_para_regions, fast_mode, data_type, repeat_region_list, result_queue):
    result_list = []
    for i in range(process_id, len(repeat_region_list), num_para_regions):
        result_list.append(quantify1repeat(f"Process {process_id:02}", 1, fast_mode, data_type, repeat_region_list[i]))
    result_queue.put(result_list)


def region_from_synth(reg):
    rr = RepeatRegion()
    rr.left_anchor_seq, rr.right_anchor_seq = reg.left_anchor_seq, reg.right_anchor_seq
    rr.left_anchor_len, rr.right_anchor_len = len(reg.left_anchor_seq), len(reg.right_anchor_seq)
    rr.repeat_unit_seq = reg.repeat_unit_seq
    for name, core, dist in zip(reg.read_names, reg.core_seqs, reg.dist_between_anchors):
        rd = Read()
        rd.read_name, rd.dist_between_anchors = name, dist
        rr.read_dict[name] = rd
        rr.read_core_seq_dict[name] = core
    return rr
