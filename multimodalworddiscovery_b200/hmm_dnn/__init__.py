"""Host-side mirrors of the reference's hmm_dnn/ word-discoverer modules."""
