class JaxNegativeBinomialMeanDisp:  # decoder only; never constructed on the distance path
    def __init__(self, *a, **k):
        raise RuntimeError("decoder is not on the local-sample-distance path")
