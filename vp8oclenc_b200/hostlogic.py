"""Host-side bookkeeping of the reference encoder that callers of the frame-level engine need:
the reference-frame rotation flags of main() and the per-segment quantiser / loop-filter table
of prepare_segments_data().  Pure Python, mirrors src/vp8enc.cpp."""
import numpy as np

_DC_Q = [4, 5, 6, 7, 8, 9, 10, 10, 11, 12, 13, 14, 15, 16, 17, 17, 18, 19, 20, 20, 21, 21, 22, 22, 23, 23, 24, 25,
         25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 46, 47, 48,
         49, 50, 51, 52, 53, 54, 55, 56, 57, 58, 59, 60, 61, 62, 63, 64, 65, 66, 67, 68, 69, 70, 71, 72, 73, 74,
         75, 76, 76, 77, 78, 79, 80, 81, 82, 83, 84, 85, 86, 87, 88, 89, 91, 93, 95, 96, 98, 100, 101, 102, 104,
         106, 108, 110, 112, 114, 116, 118, 122, 124, 126, 128, 130, 132, 134, 136, 138, 140, 143, 145, 148, 151,
         154, 157]


_AC_Q = [4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33,
         34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49, 50, 51, 52, 53, 54, 55, 56, 57, 58, 60, 62, 64,
         66, 68, 70, 72, 74, 76, 78, 80, 82, 84, 86, 88, 90, 92, 94, 96, 98, 100, 102, 104, 106, 108, 110, 112, 114,
         116, 119, 122, 125, 128, 131, 134, 137, 140, 143, 146, 149, 152, 155, 158, 161, 164, 167, 170, 173, 177, 181,
         185, 189, 193, 197, 201, 205, 209, 213, 217, 221, 225, 229, 234, 239, 245, 249, 254, 259, 264, 269, 274, 279,
         284]


def intra_quants(sd):
    """(y_dc_q, y_ac_q, uv_dc_q, uv_ac_q) of segment 0 as prepare_segments_data() derives frames.y_dc_q ... from the
    segment table (src/vp8enc.cpp:160-181); what intra_transform() quantises a key frame with"""
    cl = lambda v: max(0, min(127, int(v)))  # noqa: E731
    i = int(sd[0, 0])
    return (_DC_Q[cl(i + sd[0, 1])], _AC_Q[cl(i)], min(_DC_Q[cl(i + sd[0, 4])], 132), _AC_Q[cl(i + sd[0, 5])])


def make_segment_data(qi=(24, 24, 24, 24), key=False, lf_level=None, sharpness=0, reductor=4):
    """int32 [4][11] segment_data (src/vp8enc.h:80-92) filled like prepare_segments_data()
    (src/vp8enc.cpp:129-221) does for an inter frame; the loop-filter level is
    y_dc_q / reductor unless given explicitly."""
    sd = np.zeros((4, 11), np.int32)
    for s in range(4):
        sd[s, 0] = qi[s]
    sd[0, 1] = 15
    sd[0, 4] = 0 if key else -15
    sd[0, 5] = 0 if key else -15
    for s in range(4):
        lvl = lf_level[s] if lf_level is not None else min(63, _DC_Q[min(127, qi[s] + 15)] // reductor)
        il = lvl
        if sharpness:
            il >>= 2 if sharpness > 4 else 1
            il = min(il, 9 - sharpness)
        il = il or 1
        sd[s, 6] = lvl
        sd[s, 7] = (lvl + 2) * 2 + il
        sd[s, 8] = lvl * 2 + il
        sd[s, 9] = il
        sd[s, 10] = 0 if key else (3 if lvl >= 40 else 2 if lvl >= 20 else 1 if lvl >= 15 else 0)
    return sd


class HostState:
    """frame-type and reference bookkeeping of main() (src/vp8enc.cpp:340-374) and
    intra_transform() (src/intra_part.h:1091-1098): which frames are key / golden / altref and
    therefore which references the next inter frame may search (src/inter_part.h:35-50,103-104)."""

    def __init__(self, gop, altref_range):
        self.gop, self.altref_range = gop, altref_range
        self.until_key, self.until_altref = 1, 2
        self.n = 0
        self.golden_no = self.altref_no = -1
        self.cur_key = self.cur_golden = self.cur_altref = 0

    def next_frame(self):
        self.prev_key, self.prev_golden, self.prev_altref = self.cur_key, self.cur_golden, self.cur_altref
        self.until_key -= 1
        self.until_altref -= 1
        self.cur_key = int(self.until_key < 1)
        self.cur_golden = self.cur_key
        self.cur_altref = int(self.until_altref < 1 or self.cur_key)
        if self.until_altref < 1 or self.cur_key:
            self.until_altref = self.altref_range
        if self.cur_golden:
            self.golden_no = self.n
        if self.cur_altref:
            self.altref_no = self.n
        if self.cur_key:
            self.until_key = self.gop
        st = dict(n=self.n, key=self.cur_key, prev_golden=self.prev_golden, prev_altref=self.prev_altref,
                  altref_differs=int(self.altref_no != self.golden_no))
        self.n += 1
        return st
