"""Test-side FLAC ENCODER (test infrastructure, never imported by the product): writes streams
by the format definition (RFC 9639) so that every decoder branch of ppgs_b200/csrc/flac.cu can
be driven on seeded signals — CONSTANT / VERBATIM / FIXED 0-4 / LPC 1-32 subframes, both Rice
methods with any partition order, escaped (raw) partitions, wasted bits, independent / left-side
/ right-side / mid-side stereo, table and explicit block-size and sample-rate codes, fixed and
variable blocking, metadata blocks, ID3 wrappers, STREAMINFO MD5.  No FLAC tool or file exists in
this image; the external pin is the RFC's own example stream (tests/test_flac.py)."""
import hashlib

import numpy as np

BLOCK_CODES = {192: 1, 576: 2, 1152: 3, 2304: 4, 4608: 5, 256: 8, 512: 9, 1024: 10, 2048: 11,
               4096: 12, 8192: 13, 16384: 14, 32768: 15}
RATE_CODES = {88200: 1, 176400: 2, 192000: 3, 8000: 4, 16000: 5, 22050: 6, 24000: 7, 32000: 8,
              44100: 9, 48000: 10, 96000: 11}
SIZE_CODES = {8: 1, 12: 2, 16: 4, 20: 5, 24: 6, 32: 7}


def crc(data, width, poly):
    value, top, mask = 0, 1 << (width - 1), (1 << width) - 1
    for byte in data:
        value ^= byte << (width - 8)
        for _ in range(8):
            value = ((value << 1) ^ poly) & mask if value & top else (value << 1) & mask
    return value


def ubits(value, count):
    return format(value, '0%db' % count) if count else ''


def sbits(value, count):
    return ubits(value & ((1 << count) - 1), count) if count else ''


def coded_number(value):
    if value < 0x80:
        return bytes([value])
    for length, limit in ((2, 1 << 11), (3, 1 << 16), (4, 1 << 21), (5, 1 << 26), (6, 1 << 31), (7, 1 << 36)):
        if value < limit:
            out = []
            for _ in range(length - 1):
                out.append(0x80 | (value & 0x3f))
                value >>= 6
            lead = (0xff << (8 - length)) & 0xff
            return bytes([lead | value] + out[::-1])
    raise ValueError('number too large')


def rice_bits(residual, k):
    folded = np.where(residual >= 0, 2 * residual, -2 * residual - 1)
    return int(np.sum(folded >> k)) + len(residual) * (1 + k)


def residual_bits(residual, order, block, partition_order, method, escape):
    """method 0 / 1: 4- / 5-bit Rice parameters; escape: set of partitions stored raw."""
    param_bits, escape_code = (4, 15) if method == 0 else (5, 31)
    while partition_order and (block % (1 << partition_order) or (block >> partition_order) <= order):
        partition_order -= 1                    # short last frame: what a real encoder does too
    out = [ubits(method, 2), ubits(partition_order, 4)]
    size = block >> partition_order
    start = 0
    for part in range(1 << partition_order):
        count = size - (order if part == 0 else 0)
        chunk = residual[start:start + count]
        start += count
        if part in escape:
            need = 0 if not len(chunk) or not np.any(chunk) else int(max(
                int(chunk.max()).bit_length(), int(-chunk.min() - 1).bit_length() if chunk.min() < 0 else 0)) + 1
            out += [ubits(escape_code, param_bits), ubits(need, 5)]
            out += [sbits(int(v), need) for v in chunk]
            continue
        k = min(range(escape_code), key=lambda kk: rice_bits(chunk, kk)) if len(chunk) else 0
        out.append(ubits(k, param_bits))
        for v in chunk.tolist():
            folded = 2 * v if v >= 0 else -2 * v - 1
            out.append('0' * (folded >> k) + '1' + ubits(folded & ((1 << k) - 1), k))
    return ''.join(out)


def lpc_coefficients(signal, order, precision):
    x = signal.astype(np.float64)
    rows = np.stack([x[order - 1 - j:len(x) - 1 - j] for j in range(order)], axis=1)
    solution = np.linalg.lstsq(rows, x[order:], rcond=None)[0] if len(x) > 2 * order else np.zeros(order)
    peak = max(float(np.abs(solution).max()), 1e-9)
    shift = int(np.clip(precision - 1 - int(np.ceil(np.log2(peak) + 1e-9)), 0, 15))
    limit = (1 << (precision - 1)) - 1
    quantised = np.clip(np.round(solution * (1 << shift)), -limit - 1, limit).astype(np.int64)
    return quantised, shift


def subframe_bits(signal, bits, spec):
    """signal: int64 (block,); spec: dict(kind, order, precision, partition_order, method, escape,
    wasted).  Returns the subframe as a bit string."""
    block = len(signal)
    wasted = 0
    if spec.get('wasted') and np.any(signal):
        while not np.any(signal & ((1 << (wasted + 1)) - 1)):
            wasted += 1
    head = '0'
    signal = signal >> wasted
    bits -= wasted
    tail = ('1' + '0' * (wasted - 1) + '1') if wasted else '0'
    kind = spec['kind']
    partition = dict(partition_order=spec.get('partition_order', 0), method=spec.get('method', 0),
                     escape=spec.get('escape', ()))
    if kind == 'constant':
        assert np.all(signal == signal[0])
        return head + '000000' + tail + sbits(int(signal[0]), bits)
    if kind == 'verbatim':
        return head + '000001' + tail + ''.join(sbits(int(v), bits) for v in signal)
    order = spec['order']
    warm = ''.join(sbits(int(v), bits) for v in signal[:order])
    if kind == 'fixed':
        residual = signal.copy()
        for _ in range(order):
            residual = np.concatenate([residual[:1] * 0, np.diff(residual)])
        return (head + ubits(8 + order, 6) + tail + warm +
                residual_bits(residual[order:], order, block, **partition))
    assert kind == 'lpc'
    precision = spec.get('precision', 12)
    coefficient, shift = lpc_coefficients(signal, order, precision)
    prediction = np.zeros(block - order, dtype=np.int64)
    for j in range(order):
        prediction += coefficient[j] * signal[order - 1 - j:block - 1 - j]
    residual = signal[order:] - (prediction >> shift)
    return (head + ubits(32 + order - 1, 6) + tail + warm + ubits(precision - 1, 4) + sbits(shift, 5) +
            ''.join(sbits(int(c), precision) for c in coefficient) +
            residual_bits(residual, order, block, **partition))


def frame_bytes(samples, bits, rate, number, variable, stereo, specs, explicit_block=False,
                explicit_rate=None, streaminfo_size=False):
    """samples: int64 (channels, block)."""
    channels, block = samples.shape
    head = bytearray([0xff, 0xf9 if variable else 0xf8])
    block_code = BLOCK_CODES.get(block)
    if block_code is None or explicit_block:
        block_code = 6 if block <= 256 else 7
    if explicit_rate is None:
        rate_code = RATE_CODES.get(rate, 0)
    else:
        rate_code = explicit_rate
    head.append((block_code << 4) | rate_code)
    assignment = {'independent': channels - 1, 'left_side': 8, 'right_side': 9, 'mid_side': 10}[stereo]
    head.append((assignment << 4) | ((0 if streaminfo_size else SIZE_CODES[bits]) << 1))
    head += coded_number(number)
    if block_code == 6:
        head.append(block - 1)
    elif block_code == 7:
        head += (block - 1).to_bytes(2, 'big')
    if rate_code == 12:
        head.append(rate // 1000)
    elif rate_code == 13:
        head += rate.to_bytes(2, 'big')
    elif rate_code == 14:
        head += (rate // 10).to_bytes(2, 'big')
    head.append(crc(head, 8, 0x07))
    if stereo == 'independent':
        coded = [(samples[c], bits) for c in range(channels)]
    else:
        left, right = samples[0], samples[1]
        side = left - right
        coded = {'left_side': [(left, bits), (side, bits + 1)],
                 'right_side': [(side, bits + 1), (right, bits)],
                 'mid_side': [((left + right) >> 1, bits), (side, bits + 1)]}[stereo]
    body = ''.join(subframe_bits(signal, width, specs[c % len(specs)]) for c, (signal, width) in enumerate(coded))
    body += '0' * (-len(body) % 8)
    frame = bytes(head) + (int(body, 2).to_bytes(len(body) // 8, 'big') if body else b'')
    return frame + crc(frame, 16, 0x8005).to_bytes(2, 'big')


def encode(samples, rate, bits, blocks=4096, specs=None, stereo='independent', variable=False,
           md5=True, total_known=True, extra_metadata=True, id3=False, **frame_options):
    """samples: integer array (channels, n) within the `bits` range.  blocks: one size or a list
    of sizes cycled over the frames (a list with different sizes needs variable=True unless only
    the last frame is short).  specs: callable(frame_index) -> list of per-channel subframe specs,
    or a list of specs used for every frame."""
    samples = np.asarray(samples, dtype=np.int64)
    channels, total = samples.shape
    sizes = [blocks] if isinstance(blocks, int) else list(blocks)
    if specs is None:
        specs = [dict(kind='fixed', order=2)]
    frames, start, index = [], 0, 0
    used = []
    while start < total:
        block = min(sizes[index % len(sizes)], total - start)
        frame_specs = specs(index) if callable(specs) else specs
        frames.append(frame_bytes(samples[:, start:start + block], bits, rate, start if variable else index,
                                  variable, stereo, frame_specs, **frame_options))
        used.append(block)
        start += block
        index += 1
    width = (bits + 7) // 8
    interleaved = samples.T.reshape(-1)
    pcm = b''.join(int(v).to_bytes(width, 'little', signed=True) for v in interleaved.tolist())
    digest = hashlib.md5(pcm).digest() if md5 else bytes(16)
    max_block = max(sizes) if not variable else max(used)
    min_block = max_block if not variable else min(used)
    info = bytearray()
    info += min_block.to_bytes(2, 'big') + max_block.to_bytes(2, 'big')
    info += min(map(len, frames)).to_bytes(3, 'big') + max(map(len, frames)).to_bytes(3, 'big')
    packed = (rate << 44) | ((channels - 1) << 41) | ((bits - 1) << 36) | (total if total_known else 0)
    info += packed.to_bytes(8, 'big') + digest
    out = bytearray()
    if id3:
        out += b'ID3\x04\x00\x00' + bytes([0, 0, 0, 20]) + bytes(20)
    out += b'fLaC'
    out += bytes([0x00 if extra_metadata else 0x80]) + len(info).to_bytes(3, 'big') + info
    if extra_metadata:
        comment = b'\x04\x00\x00\x00test' + b'\x00\x00\x00\x00'
        out += bytes([0x04]) + len(comment).to_bytes(3, 'big') + comment
        out += bytes([0x81]) + (37).to_bytes(3, 'big') + bytes(37)          # PADDING, last
    for frame in frames:
        out += frame
    return bytes(out)
