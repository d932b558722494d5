"""Output channel order of the posteriorgram (ppgs/phonemes.py:10-50)."""

SILENCE = '<silent>'   # pypar.SILENCE

PHONEMES = [
    'aa', 'ae', 'ah', 'ao', 'aw', 'ay', 'b', 'ch', 'd', 'dh', 'eh', 'er', 'ey',
    'f', 'g', 'hh', 'ih', 'iy', 'jh', 'k', 'l', 'm', 'n', 'ng', 'ow', 'oy', 'p',
    'r', 's', 'sh', 't', 'th', 'uh', 'uw', 'v', 'w', 'y', 'z', 'zh', SILENCE]

PHONEME_TO_INDEX_MAPPING = {phone: i for i, phone in enumerate(PHONEMES)}
