/* stand-in: the reference includes this header in TranslationAdditionCoefficients.cpp but uses nothing from it */
