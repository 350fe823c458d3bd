# Declarations for `from pysam.libcalignmentfile cimport AlignmentFile, AlignedSegment`
# (tiddit/tiddit_signal.pyx:7).  TEST INFRASTRUCTURE ONLY: our own stand-in, see libcalignmentfile.pyx.

cdef class AlignedSegment:
    cdef public object query_name
    cdef public object query_sequence
    cdef public long flag
    cdef public long reference_id
    cdef public long reference_start
    cdef public long mapping_quality
    cdef public long next_reference_id
    cdef public long next_reference_start
    cdef public long template_length
    cdef public object _cigar
    cdef public object _tags
    cdef public object _refs


cdef class AlignmentFile:
    cdef public object header
    cdef public object references
    cdef public object lengths
    cdef public object filename
    cdef object _fh
