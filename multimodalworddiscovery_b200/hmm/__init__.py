"""Host-side mirrors of the reference's hmm/ word-discoverer modules."""
