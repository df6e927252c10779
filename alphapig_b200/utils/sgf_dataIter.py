"""SGF game records -> move-index sequences, the reader side of the supervised bootstrap
(reference utils/sgf_dataIter.py:11-66).  Same function names, arguments and result dict.

File convention of the reference data set: ``<id>_<Black|White|blank|white>..._.sgf`` -- the winner is
read from the five characters after the first underscore (:58-61); the body is
``...SZ[15]<sep>;B[hh];W[ii];...<4 trailing chars>`` and every move is the two letters after ``B[`` / ``W[``,
row letter first, index = row*15 + col (:33-36)."""
import os

LETTER_NUM = list("abcdefghijklmno")
BIG_LETTER_NUM = list("ABCDEFGHIJKLMNO")
NUM_LIST = list(range(15))
seq_lookup = dict(zip(LETTER_NUM, NUM_LIST))
num2char_lookup = dict(zip(NUM_LIST, BIG_LETTER_NUM))


def get_files_as_list(data_dir):
    """*.sgf files directly under data_dir (:22-26)"""
    return [f for f in os.listdir(data_dir) if f.endswith('.sgf') and os.path.isfile(os.path.join(data_dir, f))]


def content_to_order(sequence):
    """'B[hh];W[ii]...' -> (['hh', 'ii', ...], [112, 128, ...]) (:28-37)"""
    seq_list = [item[2:4] for item in sequence.split(';')]
    seq_num_list = [seq_lookup[item[0]] * 15 + seq_lookup[item[1]] for item in seq_list]
    return seq_list, seq_num_list


def num2char(order_):
    """move index -> 'HH'-style coordinate (:40-44; integer division made explicit for Python 3)"""
    return '%s%s' % (num2char_lookup[order_ // 15], num2char_lookup[order_ % 15])


def winner_from_name(file_name):
    tag = file_name[file_name.index('_') + 1:file_name.index('_') + 6]
    if tag in ('Blank', 'blank'):  # sic: the data set spells Black "Blank"
        return 1
    if tag in ('White', 'white'):
        return 2
    raise ValueError('file: %s carries no winner tag' % file_name)


def get_data_from_files(file_name, data_dir):
    """One record -> {'winner', 'seq_list', 'seq_num_list', 'file_name'} (:45-66)"""
    assert file_name.endswith('.sgf'), 'file: %s is not an SGF file' % file_name
    with open(os.path.join(data_dir, file_name)) as f:
        p = f.read()
    sequence = p[p.index('SZ[15]') + 7:-4]
    seq_list, seq_num_list = content_to_order(sequence)
    return {'winner': winner_from_name(file_name), 'seq_list': seq_list, 'seq_num_list': seq_num_list,
            'file_name': file_name}
